"""Streaming latency (BASELINE.json configs[3]): batch-1, 20 ms frames (and the agent's default
0.1 s chunks), tokenize_audio + detokenize_audio per step with the 2.0 s rolling context, through
the public AudioTokenizer API.  Wall clock around the Python call is the reference's own definition
(realtime_agent_profiler.py:18-38); CUDA-event time is reported beside it."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rca_b200_loader  # noqa: F401
import realtime_codec_agent_b200 as pkg


def run(tok, chunk_secs, steps, warm):
    n = int(chunk_secs * 16000)
    wav = pkg.synth_audio((steps + warm + 5) * n, seed=99).numpy()
    tok.reset_context()
    enc_ms, dec_ms, enc_ev, dec_ev = [], [], [], []
    preroll = 320
    for i in range(steps + warm):
        chunk = wav[i * n:(i + 1) * n]
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        t0 = time.perf_counter()
        e0.record()
        s = tok.tokenize_audio(chunk)
        e1.record()
        t1 = time.perf_counter()
        (sr, out), hang, preroll_left = tok.detokenize_audio(s, preroll_samples=preroll)
        e2.record()
        t2 = time.perf_counter()
        torch.cuda.synchronize()
        if i >= warm:
            enc_ms.append((t1 - t0) * 1e3); dec_ms.append((t2 - t1) * 1e3)
            enc_ev.append(e0.elapsed_time(e1)); dec_ev.append(e1.elapsed_time(e2))
    q = lambda a: {"p50": float(np.percentile(a, 50)), "p90": float(np.percentile(a, 90)), "p99": float(np.percentile(a, 99))}
    frames = max(1, int(round(chunk_secs * 50)))
    return {"chunk_secs": chunk_secs, "steps": steps, "encode_wall_ms": q(enc_ms), "decode_wall_ms": q(dec_ms),
            "encode_cuda_ms": q(enc_ev), "decode_cuda_ms": q(dec_ev),
            "decode_ms_per_frame_p50": float(np.percentile(dec_ms, 50)) / frames}


def run_emit(tok, steps, warm, chunk_secs=0.1, target_rms=0.05):
    """The agent's output step (realtime_agent_v2.py:556-579) on 0.1 s chunks: OutputChunkEmitter.emit = decoder +
    pad_or_trim + normalize_audio_rms + smooth_join in ONE engine call, wall clock around the Python call."""
    n = int(chunk_secs * 16000)
    wav = pkg.synth_audio((steps + warm + 5) * n, seed=98).numpy()
    tok.reset_context()
    strings = [tok.tokenize_audio(wav[i * n:(i + 1) * n]) for i in range(steps + warm)]
    tok.reset_context()
    em = pkg.OutputChunkEmitter(tok, chunk_secs, 0.02, target_rms)
    ms = []
    for i, s in enumerate(strings):
        t0 = time.perf_counter()
        out = em.emit(s)
        t1 = time.perf_counter()
        assert out.shape == (n,)
        if i >= warm:
            ms.append((t1 - t0) * 1e3)
    return {"chunk_secs": chunk_secs, "steps": steps, "target_volume_rms": target_rms,
            "emit_wall_ms": {"p50": float(np.percentile(ms, 50)), "p90": float(np.percentile(ms, 90)), "p99": float(np.percentile(ms, 99))}}


def main(steps=2000):
    spec = pkg.DEFAULT_SPEC
    gen = pkg.B200Generator(spec, pkg.init_random_weights(spec, seed=0), device="cuda")
    tok = pkg.AudioTokenizer(codec_model=gen, device="cuda")
    out = {"frame_20ms": run(tok, 0.02, steps, 120), "chunk_100ms": run(tok, 0.1, max(200, steps // 4), 30),
           "emit_chain_100ms": run_emit(tok, max(200, steps // 4), 30)}
    print(json.dumps(out))
    return out


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 2000)
