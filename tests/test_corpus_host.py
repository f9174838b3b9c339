"""CPU: offline-corpus host logic — window planning reproduces chunked_tokenize_audio exactly
(including warm-up windows and a ragged last chunk), LPT sharding, and the manifest all_gather over
a world_size-2 gloo group."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import realtime_codec_agent_b200 as pkg
from oracle.magicodec_oracle import OracleGenerator
from realtime_codec_agent_b200 import corpus
from tests.fake_gen import OracleBackedGen


@pytest.fixture(scope="module")
def oracle():
    return OracleGenerator(pkg.TINY_SPEC, pkg.init_random_weights(pkg.TINY_SPEC, seed=0))


@pytest.mark.parametrize("n_samples", [40000 + 700, 40000 + 100, 1600 * 7, 500])
def test_encode_streams_equals_chunked_tokenize(oracle, n_samples):
    wav = pkg.synth_audio(n_samples, file_id=11)
    tok = pkg.AudioTokenizer(codec_model=oracle, device="cpu")
    ref = tok.chunked_tokenize_audio(wav.numpy(), 0.1)
    ref_codes = np.array([ord(c) - tok.unicode_offset for c in ref])
    gen = OracleBackedGen(oracle)
    got = corpus.encode_streams(gen, [wav], 0.1, 2.0, batch_size=8, causal_warmup=False)[0].numpy()
    assert np.array_equal(got, ref_codes)
    # warm-up chunks taken from ONE prefix window (causality): on the fp32 CPU oracle a different sequence length may
    # move last bits, so frames are compared where the oracle's own top-2 margin is clear
    fast = corpus.encode_streams(OracleBackedGen(oracle), [wav], 0.1, 2.0, batch_size=8)[0].numpy()
    assert fast.shape == ref_codes.shape
    with torch.no_grad():
        _, idx, margin = oracle.quantizer.inference(oracle.encoder(oracle.pad_audio(wav[None, :32000])), return_margin=True)
    n_warm = min(len(ref_codes), 95)
    clear = margin[0, :n_warm].numpy() > 1e-3
    assert np.array_equal(fast[:n_warm][clear], ref_codes[:n_warm][clear]) and clear.mean() > 0.9
    assert np.array_equal(fast[n_warm:], ref_codes[n_warm:])


def test_streams_are_batched_across_files(oracle):
    gen = OracleBackedGen(oracle)
    a, b = pkg.synth_audio(36800, file_id=1), pkg.synth_audio(36800, file_id=2)
    both = corpus.encode_streams(gen, [a, b], 0.1, 2.0, batch_size=16, causal_warmup=False)
    assert (2, 1600) in gen.calls and (2, 30400) in gen.calls          # warm-up windows of both files share launches
    solo = corpus.encode_streams(OracleBackedGen(oracle), [b], 0.1, 2.0, batch_size=16, causal_warmup=False)[0]
    assert torch.equal(both[1], solo)
    gen2 = OracleBackedGen(oracle)
    corpus.encode_streams(gen2, [a, b], 0.1, 2.0, batch_size=16)
    assert (2, 30400) in gen2.calls and (2, 1600) not in gen2.calls    # one prefix window per file replaces the 19 warm-up launches


def test_plan_stream_shapes():
    irr, steady = corpus.plan_stream(16000 * 600, 1600, 32000, 16000, 50.0, 320)
    assert len(irr) == 19 and steady == (19, 5981, 5)
    assert [w.length for w in irr] == [1600 * k for k in range(1, 20)]
    irr, steady = corpus.plan_stream(32000 + 9280, 9280, 32000, 16000, 50.0, 320)      # 0.58 s chunks: the k=29 quirk
    assert irr[0].keep == 28


def test_shard_by_duration_is_balanced_and_deterministic():
    durs = [600.0] * 6 + [30.0, 1200.0, 45.0, 900.0]
    shards = corpus.shard_by_duration(durs, 4)
    assert sorted(i for s in shards for i in s) == list(range(len(durs)))
    loads = [sum(durs[i] for i in s) for s in shards]
    assert max(loads) - min(loads) <= 600.0
    assert shards == corpus.shard_by_duration(durs, 4)


def test_manifest_all_gather_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    from tests.dist_worker import manifest_worker
    procs = [ctx.Process(target=manifest_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert results[0] == results[1]
    assert [e[0] for e in results[0]] == [0, 1, 2, 3, 4]
    assert {e[3] for e in results[0]} == {0, 1}
    assert results[0][2][1] == 1500
