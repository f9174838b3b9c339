"""Per-kernel-class share of GPU time from an ncu launch list (`--metrics gpu__time_duration.sum --csv`): the
cross-check of bench.py's `roofline.share_of_step` (ncu times are cold-cache and serialised: shares, not absolutes)."""
import csv
import re
import sys

CLASSES = [("gemm", r"gemm2?_bf16_sm100|gemm_splitk_sm100"), ("attention", r"attention_window"), ("vq", r"vq_"),
           ("elementwise", r"rmsnorm|conv_first|gather_stem|compact_rows|embed_codes|tconv_last|emit_chunk|pack_latents|zero_pads|"
                           r"pool_roll|pcm_to_f32|resample_poly")]
rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
tot, other = {c: [0.0, 0] for c, _ in CLASSES}, [0.0, 0]
for r in rows:
    name, ns = r["Kernel Name"], float(r["Metric Value"])
    for c, pat in CLASSES:
        if re.search(pat, name):
            tot[c][0] += ns; tot[c][1] += 1
            break
    else:
        other[0] += ns; other[1] += 1
engine = sum(v[0] for v in tot.values())
print(f"{len(rows)} launches; engine kernels {engine / 1e6:.2f} ms, non-engine (torch: audio synthesis, copies) {other[0] / 1e6:.2f} ms / {other[1]} launches")
for c, (ns, n) in tot.items():
    print(f"  {c:12s} {n:6d} launches  {ns / 1e6:9.2f} ms  {100 * ns / engine:5.1f} % of engine time")
