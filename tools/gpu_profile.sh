#!/bin/bash
# One gpurun call: smoke, bench.py, the ncu launch list of one batch, and `ncu --set full` over every kernel of a
# 2+2-layer cut of the default spec, summarised ON THE BOX (the .ncu-rep is too large to travel back).
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_n1.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], "roofline", {k:d["roofline"][k] for k in ("achieved","frac","share_of_step","other_classes_ms_per_step")}, "clocks", d["clocks"], "cpu", d.get("cpu_baseline",{}).get("value"))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 256 1 > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit=$?"
timeout 300 python tools/profile_kernels.py > gpurun_out/profile_kernels.log 2>&1; echo "profile_kernels exit=$?"
timeout 1200 ncu --set full --clock-control none --profile-from-start off -o gpurun_out/prof_all python tools/profile_kernels.py > gpurun_out/ncu_all.log 2>&1
echo "ncu full exit=$?"; tail -1 gpurun_out/ncu_all.log
python tools/ncu_summary.py gpurun_out/prof_all.ncu-rep gpurun_out/ncu_all_kernels --engine-profile gpurun_out/profile_kernels.log --traffic-json gpurun_out/ncu_traffic.json > gpurun_out/ncu_summary.log 2>&1; echo "summary exit=$?"
rm -f gpurun_out/prof_all.ncu-rep
du -sh gpurun_out
