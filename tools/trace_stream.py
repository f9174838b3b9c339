"""Timeline of ONE streaming call under CUDA-graph replay (needs the -DMC_TRACE build: `python -m
realtime_codec_agent_b200.build --trace`, rebuilt here if necessary).  Prints, per kernel of the replayed graph, when its
first CTA was ready (before griddepcontrol.wait), when its dependency resolved (after), and the step to the next
kernel's resolve time = that kernel's cost on the dependency chain."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rca_b200_loader  # noqa: F401
from realtime_codec_agent_b200 import build as bld

os.environ["MAGICODEC_B200_LIB"] = bld.build(trace=True)      # libmagicodec_b200_trace.so (built here if it is stale)
import realtime_codec_agent_b200 as pkg  # noqa: E402

spec = pkg.DEFAULT_SPEC
gen = pkg.B200Generator(spec, pkg.init_random_weights(spec, seed=0), device="cuda")
for opt in sys.argv[2:]:
    k, v = opt.split("=")
    gen.set_option(k, int(v))
kind = sys.argv[1] if len(sys.argv) > 1 else "decode"
sess = gen.open_stream(1, 32000)
w = pkg.synth_audio(64000).numpy()
codes = gen.encode(torch.from_numpy(w[None, :32000]).cuda()).cpu().numpy()
for i in range(130):
    sess.push_audio(w[None, i * 320:(i + 1) * 320], 1)
    sess.push_codes(codes[:, i % 100: i % 100 + 1], 640)
CAP = 1 << 16
buf = torch.zeros(CAP * 4, dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
gen._lib.mc_debug_trace(gen._handle, buf.data_ptr(), CAP)
if kind == "decode":
    sess.push_codes(codes[:, 5:6], 640)
else:
    sess.push_audio(w[None, 130 * 320:131 * 320], 1)
n = gen._lib.mc_debug_trace(gen._handle, None, 0)
rec = buf.cpu().numpy().astype(np.uint64).reshape(-1, 4)[:n]
marks = rec[rec[:, 1] >= np.uint64(100000)]                 # phase marks of instrumented kernels (mc_trace_mark)
rec = rec[rec[:, 1] < np.uint64(100000)]
n = len(rec)
grid = (rec[:, 0] >> np.uint64(32)).astype(np.int64)
blk = (rec[:, 0] & np.uint64(0xFFFFFFFF)).astype(np.int64)
bdim = rec[:, 1].astype(np.int64)
t0, t1 = rec[:, 2].astype(np.int64), rec[:, 3].astype(np.int64)
# group CTAs of one launch: same (grid, block size) and overlapping in time -> cluster by sorted resolve time
order = np.argsort(t1, kind="stable")
launches = []
for i in order:
    key = (grid[i], bdim[i])
    if launches and launches[-1]["key"] == key and len(launches[-1]["idx"]) < grid[i]:
        launches[-1]["idx"].append(i)
    else:
        launches.append({"key": key, "idx": [i]})
base = min(t1)
print(f"# {kind}: {n} CTA records, {len(launches)} launches (kernels without griddepcontrol.wait, e.g. memsets/copies, are not listed)")
print(f"{'#':>3s} {'grid':>5s} {'threads':>7s} {'ready_us':>9s} {'resolved_us':>11s} {'last_cta_resolved':>17s} {'step_us':>8s}")
prev = None
rows = []
for k, L in enumerate(launches):
    idx = L["idx"]
    ready, res, res_last = (min(t0[idx]) - base) / 1e3, (min(t1[idx]) - base) / 1e3, (max(t1[idx]) - base) / 1e3
    rows.append((k, L["key"][0], L["key"][1], ready, res, res_last))
for j, r in enumerate(rows):
    step = rows[j + 1][4] - r[4] if j + 1 < len(rows) else float("nan")
    print(f"{r[0]:3d} {r[1]:5d} {r[2]:7d} {r[3]:9.2f} {r[4]:11.2f} {r[5]:17.2f} {step:8.2f}")

if len(marks):
    # phase marks (thread 0 of CTA 0): time since that CTA's dependency resolved, for each instrumented launch
    mt = marks[:, 2].astype(np.int64)
    mp = (marks[:, 1].astype(np.int64) - 100000)
    mg = (marks[:, 0] >> np.uint64(32)).astype(np.int64)
    res0 = {}
    for i in range(len(rec)):
        if blk[i] == 0:
            res0.setdefault(int(grid[i]), []).append(int(t1[i]))
    print("# phase marks: us after CTA 0's griddepcontrol.wait returned (one line per instrumented launch)")
    order = np.argsort(mt, kind="stable")
    line, prev_t = [], None
    for i in order:
        if prev_t is not None and mt[i] - prev_t > 20000 and line:      # > 20 us apart: the next instrumented launch
            print("  " + "  ".join(line)); line = []
        cands = [t for t in res0.get(int(mg[i]), []) if t <= mt[i]]
        ref = max(cands) if cands else mt[i]
        line.append(f"{mp[i]}:{(mt[i] - ref) / 1e3:.2f}")
        prev_t = mt[i]
    if line:
        print("  " + "  ".join(line))
