"""A/B of one engine option on the streaming path, in one process on one GPU: wall-clock p50 of the C calls a session makes
(push_audio keeping 1 frame, push_codes keeping 640 samples; 2.0 s context, 20 ms frames), alternating the option value
between blocks so clock drift hits both arms.   python tools/ab_stream_option.py fuse_qkv_attn 0 1"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rca_b200_loader  # noqa: F401
import realtime_codec_agent_b200 as pkg

key, values = sys.argv[1], [int(v) for v in sys.argv[2:]]
spec = pkg.DEFAULT_SPEC
gen = pkg.B200Generator(spec, pkg.init_random_weights(spec, seed=0), device="cuda")
w = pkg.synth_audio(64000).numpy()
codes = gen.encode(pkg.synth_audio(32000, device="cuda")[None]).cpu().numpy()
res = {v: ([], []) for v in values}
for block in range(6):
    for v in values:
        gen.set_option(key, v)
        sess = gen.open_stream(1, 32000)
        for i in range(120):
            sess.push_audio(w[None, i * 320:(i + 1) * 320], 1)
            sess.push_codes(codes[:, i % 100: i % 100 + 1], 640)
        for i in range(300):
            t0 = time.perf_counter()
            sess.push_audio(w[None, (120 + i % 60) * 320:(121 + i % 60) * 320], 1)
            t1 = time.perf_counter()
            sess.push_codes(codes[:, i % 100: i % 100 + 1], 640)
            t2 = time.perf_counter()
            res[v][0].append(t1 - t0)
            res[v][1].append(t2 - t1)
        del sess
for v in values:
    e, d = res[v]
    print(f"{key}={v}: encode p50 {np.percentile(e, 50) * 1e6:.1f} us  decode p50 {np.percentile(d, 50) * 1e6:.1f} us  "
          f"(p10 {np.percentile(d, 10) * 1e6:.1f}, p90 {np.percentile(d, 90) * 1e6:.1f})")
