"""The kernels of one streaming step (batch-1, 100-frame context): one encode keeping the last frame and one decode
keeping the last 640 samples, launched directly (same kernels the session's CUDA graph replays).  Wrapped by the
ncu launch list; also prints CUDA-event time of the direct launches and of the graph replay."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rca_b200_loader  # noqa: F401
import realtime_codec_agent_b200 as pkg

spec = pkg.DEFAULT_SPEC
gen = pkg.B200Generator(spec, pkg.init_random_weights(spec, seed=0), device="cuda")
if "nosplit" not in sys.argv:
    gen.set_option("small_m_split_k", 2)      # the kernels a streaming session uses (direct launches here, for the launch list)
wav = pkg.synth_audio(32000, device="cuda")[None]
codes = gen.encode(wav)
for _ in range(3):
    gen.encode(wav, keep_last_frames=1)
    gen.decode(codes, keep_last_samples=640)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
ev[0].record()
gen.encode(wav, keep_last_frames=1)
ev[1].record()
gen.decode(codes, keep_last_samples=640)
ev[2].record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"direct launches: encode {ev[0].elapsed_time(ev[1]) * 1e3:.0f} us, decode {ev[1].elapsed_time(ev[2]) * 1e3:.0f} us")
if "graph" in sys.argv:
    sess = gen.open_stream(1, 32000)
    w = pkg.synth_audio(64000).numpy()
    for i in range(150):
        sess.push_audio(w[None, i * 320:(i + 1) * 320], 1)
    c = codes.cpu().numpy()
    for i in range(150):
        sess.push_codes(c[:, i % 100: i % 100 + 1], 640)
    import time
    te, td = [], []
    for i in range(300):
        t0 = time.perf_counter(); sess.push_audio(w[None, (150 + i % 40) * 320:(151 + i % 40) * 320], 1); t1 = time.perf_counter()
        sess.push_codes(c[:, i % 100: i % 100 + 1], 640); t2 = time.perf_counter()
        te.append(t1 - t0); td.append(t2 - t1)
    print(f"graph replay (wall, p50): encode {np.percentile(te, 50) * 1e6:.0f} us, decode {np.percentile(td, 50) * 1e6:.0f} us")
