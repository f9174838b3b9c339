"""One warm encode + decode, then ONE profiled encode (256 windows x 2.0 s, keep 5 frames) and ONE profiled
decode (64 windows x 100 frames, keep 1600 samples) between cudaProfilerStart/Stop: the command
`ncu --profile-from-start off --set full` wraps to capture every kernel of the path once."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rca_b200_loader  # noqa: F401
import realtime_codec_agent_b200 as pkg

spec = pkg.DEFAULT_SPEC
gen = pkg.B200Generator(spec, pkg.init_random_weights(spec, seed=0), device="cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
Bd = int(sys.argv[2]) if len(sys.argv) > 2 else 64
wav = pkg.synth_audio(B * 1600 + 32000, device="cuda")


def one_pass():
    codes = gen.encode(wav, keep_last_frames=5, row_stride=1600, num_windows=B, window_samples=32000)
    full = gen.encode(wav, keep_last_frames=0, row_stride=1600, num_windows=Bd, window_samples=32000)
    rec = gen.decode(full, keep_last_samples=1600)
    return codes, rec


one_pass()
torch.cuda.synchronize()
torch.cuda.profiler.start()
codes, rec = one_pass()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("checksum", int(codes.sum()), float(rec.double().abs().sum()))
