"""Condense an .ncu-rep (ncu --set full) into the per-kernel table committed under profiles/.

Runs `ncu -i REP --page raw --csv`, keeps the metrics that show which roofline a kernel sits on, and writes
  <out>.txt   one line per launch: duration, DRAM read/write bytes and % of peak, tensor-pipe %, SM %, L2 hit, regs
  <out>.json  the same as records + per-kernel-name aggregates
and, with --engine-profile LOG (tools/profile_kernels.py's ENGINE_PROFILE line of the same pass), profiles/ncu_traffic.json:
measured DRAM bytes per GEMM launch next to the algorithmic bytes (bench.py's roofline.traffic)."""
import argparse
import csv
import io
import json
import re
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "dur_us",
    "dram__bytes_read.sum": "dram_rd",
    "dram__bytes_write.sum": "dram_wr",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
    "launch__shared_mem_per_block_dynamic": "dyn_smem",
}
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3,
              "msecond": 1e3, "nsecond": 1e-3, "second": 1e6, "s": 1e6}


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("mc::", "").replace("at::native::", "")
    return name[:56]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out")
    ap.add_argument("--engine-profile")
    ap.add_argument("--traffic-json")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    lines = [l for l in raw.splitlines() if l.startswith('"')]
    rd = csv.reader(io.StringIO("\n".join(lines)))
    header = next(rd)
    units = next(rd)
    col = {h: i for i, h in enumerate(header)}
    recs = []
    for row in rd:
        r = {"kernel": short(row[col["Kernel Name"]])}
        for m, k in KEEP.items():
            if m not in col:
                continue
            v = row[col[m]].replace(",", "")
            try:
                x = float(v)
            except ValueError:
                continue
            u = units[col[m]]
            if k in ("dram_rd", "dram_wr"):
                x *= UNIT_SCALE.get(u, 1.0)
            elif k == "dur_us":
                x *= UNIT_SCALE.get(u, 1.0)
            r[k] = x
        recs.append(r)
    with open(a.out + ".txt", "w") as f:
        f.write(f"# {len(recs)} launches from {a.rep} (ncu --set full --clock-control none); bytes are DRAM bytes of that launch\n")
        f.write(f"{'kernel':56s} {'grid':>7s} {'blk':>4s} {'regs':>4s} {'dur_us':>8s} {'dramRd_MB':>9s} {'dramWr_MB':>9s} {'GB/s':>7s} "
                f"{'dram%':>6s} {'tensor%':>7s} {'sm%':>6s} {'L2hit%':>6s} {'L2%':>6s}\n")
        for r in recs:
            gbs = (r.get("dram_rd", 0) + r.get("dram_wr", 0)) / max(r.get("dur_us", 1e-9), 1e-9) / 1e3
            f.write(f"{r['kernel']:56s} {int(r.get('grid', 0)):7d} {int(r.get('block', 0)):4d} {int(r.get('regs', 0)):4d} "
                    f"{r.get('dur_us', 0):8.1f} {r.get('dram_rd', 0) / 1e6:9.2f} {r.get('dram_wr', 0) / 1e6:9.2f} {gbs:7.0f} "
                    f"{r.get('dram_pct', 0):6.1f} {r.get('tensor_pct', 0):7.1f} {r.get('sm_pct', 0):6.1f} "
                    f"{r.get('l2_hit_pct', 0):6.1f} {r.get('l2_pct', 0):6.1f}\n")
    agg = {}
    for r in recs:
        g = agg.setdefault(r["kernel"], {"launches": 0, "dur_us": 0.0, "dram_bytes": 0.0})
        g["launches"] += 1
        g["dur_us"] += r.get("dur_us", 0)
        g["dram_bytes"] += r.get("dram_rd", 0) + r.get("dram_wr", 0)
    json.dump({"launches": recs, "by_kernel": agg}, open(a.out + ".json", "w"), indent=1)
    if a.engine_profile and a.traffic_json:
        eng = None
        for l in open(a.engine_profile):
            if l.startswith("ENGINE_PROFILE "):
                eng = json.loads(l[len("ENGINE_PROFILE "):])
        if eng:
            gem = [r for r in recs if "gemm" in r["kernel"] and "bf16_sm100" in r["kernel"]]
            tot = sum(r.get("dram_rd", 0) + r.get("dram_wr", 0) for r in gem)
            c = eng["classes"]["gemm"]
            json.dump({"source": f"{a.out}.txt (tools/profile_kernels.py: {eng['layers']}+{eng['layers']}-layer cut of the default spec, "
                                 f"{eng['windows']} encode windows + {eng['decode_windows']} decode windows)",
                       "gemm": {"launches": len(gem), "engine_launches": c["launches"],
                                "dram_bytes_per_launch": tot / max(1, len(gem)),
                                "algorithmic_bytes_per_launch": c["bytes"] / max(1, c["launches"])}},
                      open(a.traffic_json, "w"), indent=1)
    print(open(a.out + ".txt").read()[:6000])


if __name__ == "__main__":
    sys.exit(main())
