"""The command `ncu --profile-from-start off --set full` wraps to capture every kernel of the path once.

A 2+2-layer cut of the default spec (same widths, so every per-layer launch has its production shape):
one warm pass, then between cudaProfilerStart/Stop ONE encode of 256 windows x 2.0 s read through the
overlapping strided view (all rows kept: the M = 25 600 shapes of layers 0-4 of the real model) and ONE
decode of 64 windows x 100 codes.  Also prints the engine's own per-class algorithmic bytes / FLOPs of the
profiled pass (the denominators tools/ncu_summary.py puts next to ncu's measured DRAM traffic)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rca_b200_loader  # noqa: F401
import realtime_codec_agent_b200 as pkg

layers = int(os.environ.get("PROFILE_LAYERS", "2"))
spec = pkg.DEFAULT_SPEC.replace(enc_layers=layers, dec_layers=layers)
gen = pkg.B200Generator(spec, pkg.init_random_weights(spec, seed=0), device="cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
Bd = int(sys.argv[2]) if len(sys.argv) > 2 else 64
wav = pkg.synth_audio(B * 1600 + 32000, device="cuda")


def one_pass():
    codes = gen.encode(wav, keep_last_frames=0, row_stride=1600, num_windows=B, window_samples=32000)
    rec = gen.decode(codes[:Bd], keep_last_samples=1600)
    return codes, rec


one_pass()
torch.cuda.synchronize()
gen.profile_begin()
torch.cuda.profiler.start()
codes, rec = one_pass()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
prof = gen.profile_end()
print("ENGINE_PROFILE " + json.dumps({"windows": B, "decode_windows": Bd, "layers": layers, "classes": prof}))
print("checksum", int(codes.sum()), float(rec.double().abs().sum()))
