"""GPU (B200): edge cases of the C-ABI path — empty, one-sample, ragged and maximum-size inputs, clamps and error
returns (never a crash or a hang), and the host tokenizer's edge behaviour on the native engine against the same
calls on the CPU oracle model (lengths / exceptions; audio_tokenizer.py:67-149)."""
import numpy as np
import pytest
import torch

import realtime_codec_agent_b200 as pkg
from oracle.magicodec_oracle import OracleGenerator
from realtime_codec_agent_b200._native import McError

pytestmark = pytest.mark.gpu
MAXPOS = 512


@pytest.fixture(scope="module")
def gen():
    return pkg.B200Generator(pkg.TINY_SPEC, pkg.init_random_weights(pkg.TINY_SPEC, seed=0), device="cuda", max_positions=MAXPOS)


def test_encode_length_edges(gen):
    wav = pkg.synth_audio(320 * MAXPOS + 10, device="cuda")
    assert gen.encode(wav[None, :1]).shape == (1, 1)                      # one sample -> one (zero-padded) frame
    assert gen.encode(wav[None, :319]).shape == (1, 1)
    assert gen.encode(wav[None, :321]).shape == (1, 2)
    full = gen.encode(wav[None, :320 * MAXPOS])                           # exactly the RoPE table
    assert full.shape == (1, MAXPOS) and int(full.min()) >= 0 and int(full.max()) < gen.codebook_size
    with pytest.raises(McError):
        gen.encode(wav[None, :0])                                         # empty
    # keep_last_frames beyond the window clamps to "all"; 0 means all
    a = gen.encode(wav[None, :3200], keep_last_frames=999)
    assert torch.equal(a, gen.encode(wav[None, :3200]))
    # causality: a prefix's frames equal the longer window's first frames (what corpus.encode_streams relies on)
    assert torch.equal(gen.encode(wav[None, :32000])[:, :37], gen.encode(wav[None, :320 * 37]))
    # right zero padding == explicit zeros (pad_audio)
    ragged = wav[:1000]
    padded = torch.cat([ragged, torch.zeros(280, device="cuda")])
    assert torch.equal(gen.encode(ragged[None]), gen.encode(padded[None]))


def test_encode_batch_and_silence(gen):
    sil = torch.zeros(3, 6400, device="cuda")
    c = gen.encode(sil)
    assert c.shape == (3, 20) and torch.equal(c[0], c[1]) and torch.equal(c[1], c[2])
    big = pkg.synth_audio(1600 * 700 + 32000, device="cuda")
    many = gen.encode(big, keep_last_frames=5, row_stride=1600, num_windows=700, window_samples=32000)   # M = 70 000 rows
    assert many.shape == (700, 5)
    assert torch.equal(many[123], gen.encode(big[None, 123 * 1600: 123 * 1600 + 32000], keep_last_frames=5)[0])
    with pytest.raises(ValueError):
        gen.encode(big, row_stride=1600, num_windows=701 + 20, window_samples=32000)


def test_decode_edges(gen):
    codes = torch.randint(0, gen.codebook_size, (2, 50), device="cuda")
    full = gen.decode(codes)
    assert full.shape == (2, 16000) and torch.isfinite(full).all()
    assert gen.decode(codes[:, :1]).shape == (2, 320)                      # one frame
    assert torch.equal(gen.decode(codes, keep_last_samples=10 ** 9), full)  # clamp
    assert torch.equal(gen.decode(codes, keep_last_samples=1), full[:, -1:])
    assert torch.equal(gen.decode(codes, keep_last_samples=321), full[:, -321:])
    wild = codes.clone()
    wild[0, 0], wild[1, 3] = -5, gen.codebook_size + 17                    # out-of-range codes are clamped, not read
    lo = codes.clone(); lo[0, 0], lo[1, 3] = 0, gen.codebook_size - 1
    assert torch.equal(gen.decode(wild), gen.decode(lo))
    with pytest.raises(McError):
        gen.decode(codes[:, :0])
    # causal decoder: the first samples do not depend on later codes
    assert torch.equal(gen.decode(codes[:, :20]), full[:, :6400])


def test_stream_session_limits(gen):
    sess = gen.open_stream(2, 32000)
    with pytest.raises(McError, match="exceeds the session capacity"):
        sess.push_audio(np.zeros((2, 32001), np.float32), 1)
    out = sess.push_audio(np.zeros((2, 32000), np.float32), 0)            # a chunk as long as the context, all frames
    assert out.shape == (2, 100)
    out = sess.push_audio(np.zeros((2, 1), np.float32), 1)                # one sample
    assert out.shape == (2, 1)
    sess.reset()
    assert sess.push_audio(np.zeros((2, 320), np.float32), 0).shape == (2, 1)
    with pytest.raises(McError):
        sess.push_codes(np.zeros((2, 101), np.int64), 0)
    wav = sess.push_codes(np.zeros((2, 100), np.int64), 0)
    assert wav.shape == (2, 32000)
    with pytest.raises(McError, match="mono"):
        sess.set_emit(1600, 320, 0.0, 0.003, np.zeros(320, np.float32)); sess.push_codes_emit(np.zeros(5, np.int64))


def test_native_tokenizer_edge_calls_match_the_cpu_model():
    spec = pkg.TINY_SPEC
    w = pkg.init_random_weights(spec, seed=0)
    native = pkg.AudioTokenizer(codec_model=pkg.B200Generator(spec, w, device="cuda", max_positions=1024), device="cuda")
    cpu = pkg.AudioTokenizer(codec_model=OracleGenerator(spec, w), device="cpu")
    wav = pkg.synth_audio(8000).numpy()
    for tok in (native, cpu):
        with pytest.raises(RuntimeError):
            tok.tokenize_audio(wav[:0])                                   # empty audio, empty context: both raise
        tok.reset_context()
    for chunk in (wav[:3200], wav[:0], wav[:100], (wav[:1600] * 32767).astype(np.int16), (8000, wav[:800]),
                  np.stack([wav[:640], wav[640:1280]])):
        a, b = native.tokenize_audio(chunk), cpu.tokenize_audio(chunk)
        assert len(a) == len(b)                                           # int(secs * 50) chars; 0 -> whole context ([-0:])
    assert native.tokenize_context.shape == cpu.tokenize_context.shape
    for tok in (native, cpu):
        tok.reset_context()
        with pytest.raises(RuntimeError):
            tok.detokenize_audio("")                                      # nothing to decode: both raise
    s = cpu.tokenize_audio(wav[:3200])
    (sr_a, wa), ha, pa = native.detokenize_audio(s, preroll_samples=123)
    (sr_b, wb), hb, pb = cpu.detokenize_audio(s, preroll_samples=123)
    assert (sr_a, wa.shape, ha, pa) == (sr_b, wb.shape, hb, pb)
    (_, wa), _, pa = native.detokenize_audio("", preroll_samples=40)      # empty string with context: the [-0:] quirk
    (_, wb), _, pb = cpu.detokenize_audio("", preroll_samples=40)
    assert wa.shape == wb.shape and pa == pb


def test_two_tokenizers_share_one_model_object(gen):
    """clone_for_self_play hands the SAME model object to a second AudioTokenizer (realtime_agent_resources.py:46): two
    tokenizers with different chunk sizes and channel counts interleave calls on one engine handle (shared workspace,
    per-session CUDA graphs, workspace growth in between) and each must behave as if it were alone."""
    wav = pkg.synth_audio(16000 * 4, file_id=9).numpy()
    wav2 = np.stack([wav, pkg.synth_audio(16000 * 4, file_id=10).numpy()])

    def run_alone(channels, n, audio):
        tok = pkg.AudioTokenizer(codec_model=gen, num_channels=channels, device="cuda")
        out = []
        for s0 in range(0, 16000 * 3, n):
            s = tok.tokenize_audio(audio[..., s0:s0 + n])
            (_, w), _, _ = tok.detokenize_audio(s, preroll_samples=320)
            out.append((s, w.copy()))
        return out

    alone_a, alone_b = run_alone(1, 320, wav), run_alone(2, 1600, wav2)
    a = pkg.AudioTokenizer(codec_model=gen, num_channels=1, device="cuda")
    b = pkg.AudioTokenizer(codec_model=gen, num_channels=2, device="cuda")
    got_a, got_b = [], []
    ia = ib = 0
    while ia < len(alone_a) or ib < len(alone_b):
        for _ in range(5):                                   # five 20 ms frames of A per 0.1 s chunk of B
            if ia < len(alone_a):
                s = a.tokenize_audio(wav[ia * 320:(ia + 1) * 320])
                (_, w), _, _ = a.detokenize_audio(s, preroll_samples=320)
                got_a.append((s, w.copy())); ia += 1
        if ib < len(alone_b):
            s = b.tokenize_audio(wav2[:, ib * 1600:(ib + 1) * 1600])
            (_, w), _, _ = b.detokenize_audio(s, preroll_samples=320)
            got_b.append((s, w.copy())); ib += 1
        if ib == 7:                                          # a big one-shot call grows the shared workspace mid-stream
            gen.encode(torch.zeros(64, 32000, device="cuda"))
    for got, alone in ((got_a, alone_a), (got_b, alone_b)):
        assert len(got) == len(alone)
        for (s1, w1), (s2, w2) in zip(got, alone):
            assert s1 == s2 and w1.shape == w2.shape and np.array_equal(w1, w2)


def test_inputs_longer_than_the_rope_table_grow_it():
    """The reference tokenizer takes any length in one call (run_demo.py:55,109): past max_positions the RoPE tables
    are rebuilt at the next power of two and re-registered; a live session's captured graphs re-capture."""
    g = pkg.B200Generator(pkg.TINY_SPEC, pkg.init_random_weights(pkg.TINY_SPEC, seed=0), device="cuda", max_positions=128)
    wav = pkg.synth_audio(320 * 700, device="cuda")
    sess = g.open_stream(1, 32000)
    w = wav.cpu().numpy()
    first = [sess.push_audio(w[None, i * 1600:(i + 1) * 1600], 5).copy() for i in range(25)]     # graphs captured at 128 rows
    short = g.encode(wav[None, :320 * 100])
    assert g.max_positions == 128
    long = g.encode(wav[None, :320 * 700])                                 # 700 frames > 128: grows to 1024
    assert g.max_positions == 1024 and long.shape == (1, 700)
    assert torch.equal(long[:, :100], short)                               # same angles, causal prefix
    assert torch.equal(g.encode(wav[None, :320 * 100]), short)
    rec = g.decode(long)
    assert rec.shape == (1, 700 * 320) and torch.isfinite(rec).all()
    sess.reset()
    again = [sess.push_audio(w[None, i * 1600:(i + 1) * 1600], 5).copy() for i in range(25)]     # replays re-captured graphs
    assert all(np.array_equal(a, b) for a, b in zip(first, again))
    tok = pkg.AudioTokenizer(codec_model=g, device="cuda")
    s = tok.tokenize_audio(pkg.synth_audio(16000 * 45, file_id=2).numpy())  # 45 s one-shot: 2250 frames
    assert len(s) == 2250 and g.max_positions == 4096
    (_, out), _, _ = tok.detokenize_audio(s)
    assert out.shape == (16000 * 45,)


def test_session_reseed_after_one_shot_calls_is_an_upload(gen):
    """A one-shot call longer than the session capacity goes through the stateless path; the device context is then
    re-seeded by mc_stream_load_* (no second network pass) and streaming continues exactly as if nothing happened."""
    wav = pkg.synth_audio(16000 * 8, file_id=3).numpy()
    tok = pkg.AudioTokenizer(codec_model=gen, device="cuda")
    ref = pkg.AudioTokenizer(codec_model=gen, device="cuda")
    ref._stream_session().set_graphs(False)
    launches = []
    outs, ref_outs = [], []
    for t, sink in ((tok, outs), (ref, ref_outs)):
        t.reset_context()
        sink.append(t.tokenize_audio(wav[:1600]))
        l0 = gen.launch_count
        sink.append(t.tokenize_audio(wav[1600:1600 + 48000]))              # 3 s > 2 s capacity: stateless + re-seed
        launches.append(gen.launch_count - l0)
        for i in range(6):
            sink.append(t.tokenize_audio(wav[49600 + i * 1600: 49600 + (i + 1) * 1600]))
        s_all = "".join(sink[:3])
        (_, w1), _, _ = t.detokenize_audio(s_all)                          # 3.1 s of codes > 100 frames: stateless + re-seed
        (_, w2), _, _ = t.detokenize_audio(sink[3], preroll_samples=320)
        sink.append((w1, w2))
    l0 = gen.launch_count
    gen.encode(torch.from_numpy(wav[None, 1600:49600]).cuda(), keep_last_frames=150)   # the context keeps max(n_new, 2 s) samples
    assert launches[0] == launches[1] == gen.launch_count - l0             # exactly one pass: the re-seed computes nothing
    # the same calls against windows built by hand (stateless calls with the session's kernels)
    gen.set_option("small_m_split_k", 2)
    ctx = wav[49600 + 5 * 1600 + 1600 - 32000: 49600 + 6 * 1600]
    want = gen.encode(torch.from_numpy(ctx[None]).cuda(), keep_last_frames=5)[0].cpu().numpy()
    assert [ord(c) - tok.unicode_offset for c in outs[7]] == list(want)
    assert outs[:8] == ref_outs[:8]                                        # graph replay == direct launches after the re-seed
    assert all(np.array_equal(a, b) for a, b in zip(outs[8], ref_outs[8]))
    # stateless reference for the decode that followed the re-seed of the code context
    codes = np.array([ord(c) - tok.unicode_offset for c in ("".join(outs[:3]) + outs[3])[-100:]], dtype=np.int64)
    want_w = gen.decode(torch.from_numpy(codes[None]).cuda(), keep_last_samples=1600 + 320)[0].cpu().numpy()
    gen.set_option("small_m_split_k", 1)
    assert np.array_equal(outs[8][1], want_w)


def test_threads_sharing_one_tokenizer_and_one_model(gen):
    """tts_server.py:59,158: Flask threads call tokenize_audio on ONE AudioTokenizer; realtime_agent_resources.py:41-49:
    two tokenizers on one model.  Calls are serialised per tokenizer and per engine handle: no exception, no torn
    context, and per-tokenizer results equal a single-threaded run whenever the call order per tokenizer is fixed."""
    import threading
    wav = pkg.synth_audio(16000 * 6, file_id=21).numpy()

    def sequence(tok, n_chunks, chunk):
        out = []
        for i in range(n_chunks):
            s = tok.tokenize_audio(wav[i * chunk:(i + 1) * chunk])
            (_, w), _, _ = tok.detokenize_audio(s, preroll_samples=320)
            out.append((s, w.copy()))
        return out

    ref_a = sequence(pkg.AudioTokenizer(codec_model=gen, device="cuda"), 40, 1600)
    ref_b = sequence(pkg.AudioTokenizer(codec_model=gen, device="cuda"), 120, 320)
    tok_a, tok_b = pkg.AudioTokenizer(codec_model=gen, device="cuda"), pkg.AudioTokenizer(codec_model=gen, device="cuda")
    res, errs = {}, []

    def worker(name, tok, n, chunk):
        try:
            with torch.cuda.stream(torch.cuda.Stream()):                    # each thread on its own CUDA stream
                res[name] = sequence(tok, n, chunk)
        except Exception as ex:                                             # noqa: BLE001
            errs.append(ex)

    ts = [threading.Thread(target=worker, args=("a", tok_a, 40, 1600)), threading.Thread(target=worker, args=("b", tok_b, 120, 320))]
    [t.start() for t in ts]; [t.join() for t in ts]
    assert not errs, errs
    for got, ref in ((res["a"], ref_a), (res["b"], ref_b)):
        assert all(s1 == s2 and np.array_equal(w1, w2) for (s1, w1), (s2, w2) in zip(got, ref))
    # many threads on ONE tokenizer: order is arbitrary, but every call must see a consistent context
    shared = pkg.AudioTokenizer(codec_model=gen, device="cuda")
    lens = []

    def hammer():
        try:
            for i in range(30):
                lens.append(len(shared.tokenize_audio(wav[i * 1600:(i + 1) * 1600])))
        except Exception as ex:                                             # noqa: BLE001
            errs.append(ex)

    ts = [threading.Thread(target=hammer) for _ in range(4)]
    [t.start() for t in ts]; [t.join() for t in ts]
    assert not errs and lens == [5] * 120 and shared.tokenize_context.shape == (1, 32000)


def test_look_ahead_spec_skips_the_causal_warmup_shortcut():
    """corpus.encode_streams takes warm-up chunks from one prefix window only for causal specs; with window_right > 0
    a prefix frame would see audio beyond its own window, so the per-window path must be used."""
    from realtime_codec_agent_b200 import corpus
    spec = pkg.TINY_SPEC.replace(window_left=24, window_right=8)
    g = pkg.B200Generator(spec, pkg.init_random_weights(spec, seed=0), device="cuda", max_positions=256)
    streams = [pkg.synth_audio(16000 * 3, file_id=i, device="cuda") for i in range(2)]
    fast = corpus.encode_streams(g, streams, 0.1, 2.0, batch_size=16)
    slow = corpus.encode_streams(g, streams, 0.1, 2.0, batch_size=16, causal_warmup=False)
    assert all(torch.equal(a, b) for a, b in zip(fast, slow))
    tok = pkg.AudioTokenizer(codec_model=g, device="cuda")
    s = tok.chunked_tokenize_audio(streams[0].cpu().numpy(), 0.1)
    assert [ord(c) - tok.unicode_offset for c in s] == fast[0].cpu().tolist()
    # and the shortcut WOULD have been wrong here: the prefix trick changes warm-up codes under look-ahead
    whole = g.encode(streams[0][None, :32000])[0]
    assert not torch.equal(whole[:95], fast[0][:95])


@pytest.mark.parametrize("context_secs", [0.3, 2.0, 4.0, 6.5])
@pytest.mark.parametrize("mode", [0, 2])
def test_streaming_with_other_context_lengths_equals_hand_built_windows(gen, context_secs, mode):
    """context_secs is a constructor argument of the reference (audio_tokenizer.py:15; realtime_agent_config.py sets it):
    0.3 s (15 frames), 2.0 s (one 128-row tile), 4.0 s (200 rows: two tiles, per-layer row compaction inside the few-rows
    regime) and 6.5 s (325 rows: beyond the few-rows regime).  Every streamed call must equal the stateless engine call on
    the window the reference would have built, with the session's kernels (mode 2) and the batch-invariant ones (mode 0)."""
    wav = pkg.synth_audio(16000 * 9, file_id=31).numpy()
    gen.set_option("small_m_split_k", mode)
    try:
        tok = pkg.AudioTokenizer(codec_model=gen, context_secs=context_secs, device="cuda")
        ctx_samples, ctx_frames = int(context_secs * 16000), int(context_secs * 50)
        chunk, text = 1600, ""
        for i in range(int(8.5 * 16000) // chunk):
            s = tok.tokenize_audio(wav[i * chunk:(i + 1) * chunk])
            lo = max(0, (i + 1) * chunk - max(ctx_samples, chunk))
            window = torch.from_numpy(wav[None, lo:(i + 1) * chunk]).cuda()
            want = gen.encode(window, keep_last_frames=5)[0].cpu().numpy()
            assert [ord(c) - tok.unicode_offset for c in s] == list(want), f"encode call {i}"
            text += s
            (_, w), _, _ = tok.detokenize_audio(s, preroll_samples=320)
            codes = np.array([ord(c) - tok.unicode_offset for c in text[-max(ctx_frames, 5):]], dtype=np.int64)
            want_w = gen.decode(torch.from_numpy(codes[None]).cuda(), keep_last_samples=1600 + 320)[0].cpu().numpy()
            assert np.array_equal(w, want_w), f"decode call {i}"
    finally:
        gen.set_option("small_m_split_k", 1)
