#!/bin/bash
# bench (N=1) + ncu launch list of one batch + ncu --set full on the top kernels
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_n1.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], "roofline", {k:d["roofline"][k] for k in ("achieved","frac","share_of_step","other_classes_ms_per_step")}, "clocks", d["clocks"], "cpu", d.get("cpu_baseline",{}).get("value"))
PY
timeout 300 python tools/profile_step.py > gpurun_out/profile_step.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_list.log 2>&1
tail -4 gpurun_out/profile_step.log
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"gemm2_bf16|attention_window_sm100_v2" -s 22 -c 5 -o gpurun_out/prof_top python tools/profile_step.py > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit=$?"; tail -2 gpurun_out/ncu_full.log
