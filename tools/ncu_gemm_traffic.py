"""profiles/ncu_traffic.json: DRAM bytes per launch of the four block GEMMs AT THE BENCH SHAPE, next to their algorithmic bytes.

Input: the <out>.json that tools/ncu_summary.py writes from
    ncu --set full --clock-control none --profile-from-start off -k regex:gemm2_bf16 -c 16 python tools/profile_step.py 1024 1
i.e. the CTA-pair GEMM launches of the first layers of ONE 1 024-window batch of the 8-layer default spec (M = 102 400 rows,
what bench.py issues).  The launches are told apart by their specialised epilogue and their order inside a layer:
QKV = <256, EPI_ROPE_BF16>, W1 = <256, EPI_GELU_BF16>, and of the two <256, EPI_RESID_F32> launches the one that follows QKV's
attention is Wo and the one that follows W1 is W2."""
import json
import sys

M, D, F = 102400, 1024, 4096
ALGO = {  # bytes one launch must move: A (bf16) + W (bf16) + output (bf16, or fp32 read-modify-write for the residual)
    "QKV": M * D * 2 + 3 * D * D * 2 + M * 3 * D * 2,
    "Wo": M * D * 2 + D * D * 2 + M * D * 8,
    "W1": M * D * 2 + F * D * 2 + M * F * 2,
    "W2": M * F * 2 + D * F * 2 + M * D * 8,
}


def main():
    src, out = sys.argv[1], sys.argv[2]
    recs = [r for r in json.load(open(src))["launches"] if "gemm2_bf16_sm100_kernel" in r["kernel"]]
    per = {k: [] for k in ALGO}
    prev = None
    for r in recs:
        k = r["kernel"].replace(" ", "")
        if "<256,1>" in k:
            name = "QKV"
        elif "<256,2>" in k:
            name = "W1"
        elif "<256,3>" in k:
            name = "Wo" if prev == "QKV" else "W2"
        else:
            continue
        prev = name
        if int(r.get("grid", 0)) == 148:
            per[name].append(r)
    res = {}
    for name, rs in per.items():
        if not rs:
            continue
        n = len(rs)
        dram = sum(r.get("dram_rd", 0) + r.get("dram_wr", 0) for r in rs) / n
        res[name] = {"launches_captured": n, "dram_bytes_per_launch": dram, "algorithmic_bytes_per_launch": ALGO[name],
                     "dram_over_algorithmic": dram / ALGO[name], "dur_us": sum(r.get("dur_us", 0) for r in rs) / n,
                     "tensor_pct": sum(r.get("tensor_pct", 0) for r in rs) / n, "dram_pct": sum(r.get("dram_pct", 0) for r in rs) / n,
                     "l2_hit_pct": sum(r.get("l2_hit_pct", 0) for r in rs) / n}
    json.dump({"source": f"{src} (ncu --set full of tools/profile_step.py 1024 1: one 1 024-window batch, 8 layers, M = {M} rows — the bench shape)",
               "per_gemm": res}, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
