"""GPU (B200): rows f3 / f4 — emit_chunk_kernel and embed_distance_kernel against the golden vectors the
reference's own utils/audio_utils.py + ExternalTTSDuplexAligner produced (tests/golden/make_golden_post.py),
and the fused push_codes_emit call against the oracle chain fed with the engine's own decoder output.

Tolerances: with target_volume_rms = 0 (the reference default, realtime_agent_config.py:25) the chain is
multiplies and adds with numpy's roundings -> BIT-EXACT.  With RMS normalisation the gain comes from an fp64
device sum vs numpy's fp32 pairwise sum: relative 2e-6 (absolute 1e-7 where the crossfade cancels).  Aligner scores: relative 1e-5 (fp32 norms, fp64 means).
"""
import os

import numpy as np
import pytest
import torch

import realtime_codec_agent_b200 as pkg
from oracle import post_decode_oracle as po

pytestmark = pytest.mark.gpu
SR, CHUNK, L = 16000, 1600, 320
RMS_RTOL = 2e-6
RMS_ATOL = 1e-7      # cross-faded samples are sums of two products and may cancel; 1e-7 is -140 dB re full scale


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "golden_post.npz"))


@pytest.fixture(scope="module")
def gen():
    return pkg.B200Generator(pkg.TINY_SPEC, pkg.init_random_weights(pkg.TINY_SPEC, seed=0), device="cuda", max_positions=1024)


@pytest.mark.parametrize("tag,target", [("plain", 0.0), ("rms", 0.05)])
def test_emit_kernel_replays_reference_golden(g, gen, tag, target):
    fade_in = torch.from_numpy(g["fade_in"]).cuda()
    prev_tail = torch.zeros(L, device="cuda")
    ends = np.cumsum(g["chain_lengths"])
    history = []
    for i, (e, n) in enumerate(zip(ends, g["chain_lengths"])):
        wav = torch.from_numpy(g["chain_decoded"][e - n:e]).cuda()
        out = gen.op_emit_chunk(wav, CHUNK, L, i > 0, target, 0.003, fade_in, prev_tail).cpu().numpy()
        emitted, cross, fresh = out[:CHUNK], out[CHUNK:CHUNK + L], out[CHUNK + L:]
        want = g[f"chain_emitted_{tag}"][i]
        if target == 0.0:
            assert np.array_equal(emitted, want), f"chunk {i}"
        else:
            assert np.allclose(emitted, want, rtol=RMS_RTOL, atol=RMS_ATOL), f"chunk {i}"
        if i > 0:
            history[-1] = np.concatenate((history[-1][:CHUNK - L], cross))
        history.append(fresh)
    hist = np.concatenate(history)
    if target == 0.0:
        assert np.array_equal(hist, g[f"chain_history_{tag}"])
    else:
        assert np.allclose(hist, g[f"chain_history_{tag}"], rtol=RMS_RTOL, atol=RMS_ATOL)


def test_embed_distance_kernel_replays_reference_golden(g, gen):
    table = torch.from_numpy(g["ali_table"]).cuda()
    sil = torch.from_numpy(g["ali_silence_codes"]).cuda()[None]
    _, mean = gen.op_embed_distance(table, sil, 0, None, want_mean=True)
    assert np.allclose(mean[0].cpu().numpy(), g["ali_silence_embedding"], rtol=1e-5, atol=1e-7)
    for t, d, want in zip(g["ali_tts"], g["ali_duplex"], g["ali_scores"]):
        ids = torch.from_numpy(np.stack([t, d])).cuda()
        tts, dup = gen.op_embed_distance(table, ids, 128256, mean[0]).tolist()
        assert abs(tts / (dup + 1e-5) - want) <= 1e-5 * want


@pytest.mark.parametrize("target", [0.0, 0.05])
def test_fused_emit_call_equals_oracle_chain_on_engine_output(gen, target):
    """OutputChunkEmitter (ONE C call per chunk: decode + RMS + crossfade in one graph) vs the oracle chain
    stepping a second native tokenizer through detokenize_audio — same engine decode on both sides, so the
    difference is the post chain alone.  40 chunks: covers context fill, graph capture and replay."""
    tok_a = pkg.AudioTokenizer(codec_model=gen, device="cuda")
    tok_b = pkg.AudioTokenizer(codec_model=gen, device="cuda")
    s = tok_a.tokenize_audio(pkg.synth_audio(4 * SR, file_id=4).numpy())
    tok_a.reset_context()
    emitter = pkg.OutputChunkEmitter(tok_a, 0.1, 0.02, target)
    chain = po.OracleOutputChain(tok_b, 0.1, 0.02, target)
    for i in range(0, len(s), 5):
        got, want = emitter.emit(s[i:i + 5]), chain.step(s[i:i + 5])
        assert got.shape == want.shape == (CHUNK,)
        if target == 0.0:
            assert np.array_equal(got, want), f"chunk {i // 5}"
        else:
            assert np.allclose(got, want, rtol=RMS_RTOL, atol=RMS_ATOL), f"chunk {i // 5}"
    a, b = np.concatenate(emitter.audio_history_ch1), np.concatenate(chain.history)
    assert a.shape == b.shape and np.allclose(a, b, rtol=RMS_RTOL if target else 0.0, atol=RMS_ATOL if target else 0.0)
    assert tok_a.detokenize_context == tok_b.detokenize_context
    # reset -> first-chunk behaviour again
    emitter.reset(); tok_a.reset_context()
    chain.history = []; tok_b.reset_context()
    assert np.allclose(emitter.emit(s[:5]), chain.step(s[:5]), rtol=RMS_RTOL, atol=RMS_ATOL)


def test_native_aligner_matches_oracle_on_engine_codebook(gen):
    tok = pkg.AudioTokenizer(codec_model=gen, device="cuda")
    al = pkg.ExternalTTSDuplexAligner(tok, codec_vocab_start=1000)
    q = gen.quantizer
    table = q.codebook_proj(q.codebook.weight).cpu()          # fp32 outside autocast (get_codec_embeddings gives bf16 on GPU)
    assert table.dtype == torch.float32
    sil = tok._encode_silence(10.0)[0, 0].cpu()
    emb = po.oracle_silence_embedding(table, sil)
    assert np.allclose(al.silence_embedding.cpu().numpy(), emb.numpy(), rtol=1e-5, atol=1e-7)
    rng = np.random.default_rng(3)
    K = gen.codebook_size
    for n in (1, 5, 25, 300):
        t = (rng.integers(0, K, size=n) + 1000).tolist()
        d = (rng.integers(0, K, size=n) + 1000).tolist()
        want = po.oracle_interrupt_score(table, emb, 1000, t, d)
        assert abs(al.interrupt_score(t, d) - want) <= 1e-5 * want
