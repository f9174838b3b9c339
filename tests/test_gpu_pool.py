"""GPU (B200): multi-session batching — SessionBatcher / mc_pool_* against S independent
AudioTokenizers (string- and sample-level equality), the threaded tts_server.py calling pattern, argument errors,
and the cost of 8 sessions per launch relative to one."""
import threading
import time

import numpy as np
import pytest
import torch

import realtime_codec_agent_b200 as pkg
from realtime_codec_agent_b200._native import McError
from realtime_codec_agent_b200.session_batcher import SessionBatcher, SessionPool, ThreadedSessionBatcher

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gen():
    return pkg.B200Generator(pkg.TINY_SPEC, pkg.init_random_weights(pkg.TINY_SPEC, seed=0), device="cuda", max_positions=256)


def _audio(i, secs=5.0):
    return pkg.synth_audio(int(16000 * secs), file_id=60 + i).numpy()


@pytest.mark.parametrize("channels", [1, 2])
def test_batched_sessions_equal_independent_tokenizers(gen, channels):
    """Sessions start at different times (contexts of different length next to steady-state ones), chunk sizes differ
    between ticks; with the batch-invariant kernels every session's strings and waveforms equal its own tokenizer's."""
    S = 6
    gen.set_option("small_m_split_k", 0)
    try:
        bat = SessionBatcher(gen, num_channels=channels, max_sessions=8)
        toks = [pkg.AudioTokenizer(codec_model=gen, num_channels=channels, device="cuda") for _ in range(S)]
        wav = [np.stack([_audio(i * 2 + c) for c in range(channels)]) if channels > 1 else _audio(i) for i in range(S)]
        sids, pos = {}, [0] * S
        launches = 0
        for tick in range(34):
            n = 1600 if tick % 7 else 320                               # mostly 0.1 s chunks, sometimes a 20 ms frame
            active = [i for i in range(S) if tick >= 3 * i]              # session i joins at tick 3*i
            for i in active:
                if i not in sids:
                    sids[i] = bat.open_session()
            chunks = {sids[i]: wav[i][..., pos[i]:pos[i] + n] for i in active}
            l0 = gen.launch_count
            got = bat.tokenize_audio(chunks)
            launches = gen.launch_count - l0
            dec_in = {}
            for i in active:
                ref = toks[i].tokenize_audio(wav[i][..., pos[i]:pos[i] + n])
                assert got[sids[i]] == ref, f"tick {tick} session {i}"
                pos[i] += n
                dec_in[sids[i]] = ref
            out = bat.detokenize_audio(dec_in, preroll_samples=320)
            for i in active:
                (sr, w), hang, pre = out[sids[i]]
                (sr2, w2), hang2, pre2 = toks[i].detokenize_audio(dec_in[sids[i]], preroll_samples=320)
                assert (sr, hang, pre) == (sr2, hang2, pre2) and w.shape == w2.shape and np.array_equal(w, w2), f"tick {tick} session {i}"
        assert launches == 1 or launches < 60                            # steady state: all six sessions in ONE graph replay
        bat.close_session(sids[0])
        again = bat.open_session()                                       # a recycled slot starts from an empty context
        assert bat.tokenize_audio({again: wav[0][..., :1600]})[again] == pkg.AudioTokenizer(
            codec_model=gen, num_channels=channels, device="cuda").tokenize_audio(wav[0][..., :1600])
    finally:
        gen.set_option("small_m_split_k", 1)


def test_default_kernels_agree_with_independent_tokenizers(gen):
    """Default mode: a lone session takes the split-K few-rows kernels, a batch of 8 the tiled ones — same products,
    different fp32 summation order, so near-tie codes may differ (never the clear ones)."""
    bat = SessionBatcher(gen, max_sessions=8)
    sids = [bat.open_session() for _ in range(8)]
    toks = [pkg.AudioTokenizer(codec_model=gen, device="cuda") for _ in range(8)]
    agree = []
    for tick in range(30):
        got = bat.tokenize_audio({sids[i]: _audio(i)[tick * 1600:(tick + 1) * 1600] for i in range(8)})
        for i in range(8):
            ref = toks[i].tokenize_audio(_audio(i)[tick * 1600:(tick + 1) * 1600])
            agree += [a == b for a, b in zip(got[sids[i]], ref)]
    print(f"[pool] batched (tiled kernels) vs lone sessions (split-K kernels): {np.mean(agree):.4f} of {len(agree)} codes equal")
    assert np.mean(agree) > 0.9


def test_pool_argument_errors(gen):
    pool = SessionPool(gen, 1, 32000, 4)
    x = np.zeros((2, 1, 1600), np.float32)
    pool.push_audio([0], x[:1], 5)
    with pytest.raises(McError, match="different length"):
        pool.push_audio([0, 1], x, 5)                                    # slot 0 holds 1600 samples, slot 1 none
    with pytest.raises(McError, match="listed twice"):
        pool.push_audio([2, 2], x, 5)
    with pytest.raises(McError, match="out of range"):
        pool.push_audio([4], x[:1], 5)
    with pytest.raises(McError):
        pool.push_audio([1], np.zeros((1, 1, 32001), np.float32), 5)
    with pytest.raises(McError):
        pool.push_codes([1], np.zeros((1, 1, 101), np.int64), 0)
    assert pool.context_len(0) == (1600, 0) and pool.context_len(1) == (0, 0)
    assert pool.push_codes([1, 2, 3], np.zeros((3, 1, 5), np.int64), 0).shape == (3, 1, 1600)
    bat = SessionBatcher(gen, max_sessions=1)
    bat.open_session()
    with pytest.raises(RuntimeError, match="in use"):
        bat.open_session()


def test_threaded_batcher_serves_request_threads(gen):
    """tts_server.py:59,158: request threads tokenize their own stream chunk by chunk; the dispatcher batches whatever
    arrived in the same tick.  Per-session results equal a private tokenizer (batch-invariant kernels)."""
    gen.set_option("small_m_split_k", 0)
    bat = ThreadedSessionBatcher(gen, max_sessions=8, max_wait_ms=1.0)
    try:
        S, ticks = 5, 24
        refs = []
        for i in range(S):
            tok = pkg.AudioTokenizer(codec_model=gen, device="cuda")
            refs.append([tok.tokenize_audio(_audio(i)[t * 1600:(t + 1) * 1600]) for t in range(ticks)])
        got, errs = {}, []

        def stream(i):
            try:
                sid = bat.open_session()
                got[i] = [bat.tokenize_audio_one(sid, _audio(i)[t * 1600:(t + 1) * 1600]) for t in range(ticks)]
            except Exception as ex:                                      # noqa: BLE001
                errs.append(ex)

        l0 = gen.launch_count
        ts = [threading.Thread(target=stream, args=(i,)) for i in range(S)]
        [t.start() for t in ts]; [t.join() for t in ts]
        assert not errs, errs
        assert all(got[i] == refs[i] for i in range(S))
        print(f"[pool] {S} threads x {ticks} chunks took {gen.launch_count - l0} engine launches (graph replays count as one)")
    finally:
        bat.shutdown()
        gen.set_option("small_m_split_k", 1)


def test_a_malformed_request_fails_alone(gen):
    """One bad request in a tick (empty chunk, closed session) must not take the other streams down, and a rejected
    batch must leave every session's context untouched (validation happens before anything is pushed)."""
    gen.set_option("small_m_split_k", 0)
    bat = ThreadedSessionBatcher(gen, max_sessions=4, max_wait_ms=20.0)
    try:
        good, bad = bat.open_session(), bat.open_session()
        ref = pkg.AudioTokenizer(codec_model=gen, device="cuda")
        res = {}

        def call(name, sid, chunk):
            try:
                res[name] = bat.tokenize_audio_one(sid, chunk)
            except Exception as ex:                                      # noqa: BLE001
                res[name] = ex

        ts = [threading.Thread(target=call, args=("good", good, _audio(0)[:1600])),
              threading.Thread(target=call, args=("empty", bad, np.zeros(0, np.float32))),
              threading.Thread(target=call, args=("closed", 12345, _audio(1)[:1600]))]
        [t.start() for t in ts]; [t.join() for t in ts]
        assert isinstance(res["empty"], ValueError) and isinstance(res["closed"], KeyError)
        assert res["good"] == ref.tokenize_audio(_audio(0)[:1600])
        # detokenize: the bad string is found before any context changes
        s = res["good"]
        with pytest.raises(ValueError):
            bat.detokenize_audio({good: s, bad: ""})
        assert bat._sessions[good].detok_context == "" and bat.pool.context_len(bat._sessions[good].slot)[1] == 0
        out = bat.detokenize_audio({good: s})
        assert out[good][0][1].shape == (1600,)
    finally:
        bat.shutdown()
        gen.set_option("small_m_split_k", 1)


def test_eight_sessions_cost_less_than_two():
    """Default spec: one launch for 8 sessions (800 rows) against one session alone (100 rows)."""
    spec = pkg.DEFAULT_SPEC
    g = pkg.B200Generator(spec, pkg.init_random_weights(spec, seed=0), device="cuda")
    bat = SessionBatcher(g, max_sessions=8)
    sids = [bat.open_session() for _ in range(8)]
    wav = [_audio(i, 8.0) for i in range(8)]

    def run(active, ticks, t0):
        enc, dec = [], []
        for t in range(t0, t0 + ticks):
            a = time.perf_counter()
            s = bat.tokenize_audio({sids[i]: wav[i][t * 320:(t + 1) * 320] for i in active})
            b = time.perf_counter()
            bat.detokenize_audio(s, preroll_samples=320)
            c = time.perf_counter()
            enc.append(b - a); dec.append(c - b)
        return np.median(enc[-40:]) * 1e3, np.median(dec[-40:]) * 1e3

    run(range(8), 130, 0)                                                # fill every context (and capture the 8-wide graphs)
    e8, d8 = run(range(8), 60, 130)
    e1, d1 = run([0], 60, 190)                                           # session 0 alone (others idle)
    print(f"[pool] default spec, 20 ms frames: 8 sessions per call encode {e8:.3f} ms / decode {d8:.3f} ms; "
          f"1 session encode {e1:.3f} ms / decode {d1:.3f} ms -> x{e8 / e1:.2f} / x{d8 / d1:.2f}")
    assert e8 < 2.0 * e1 and d8 < 2.0 * d1
