#!/bin/bash
# tests -> bench -> A/B step profile -> ncu launch list -> ncu --set full (2+2-layer cut), summarised ON THE BOX
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit=$?"
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_n1.json"))
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], "roofline", {k:d["roofline"][k] for k in ("achieved","frac","share_of_step","other_classes_ms_per_step")}, "clocks", d["clocks"], "cpu", d.get("cpu_baseline",{}).get("value"), "stream", d.get("streaming",{}).get("p50_decode_ms_per_frame"))
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/bench_n1.err
timeout 600 python tools/profile_step.py > gpurun_out/profile_step.log 2>&1; echo "profile_step exit=$?"; cat gpurun_out/profile_step.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 256 1 > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit=$?"
timeout 300 python tools/profile_kernels.py > gpurun_out/profile_kernels.log 2>&1; echo "profile_kernels exit=$?"; tail -2 gpurun_out/profile_kernels.log | cut -c1-600
timeout 1200 ncu --set full --clock-control none --profile-from-start off -o gpurun_out/prof_all python tools/profile_kernels.py > gpurun_out/ncu_all.log 2>&1
echo "ncu full exit=$?"; tail -2 gpurun_out/ncu_all.log; ls -la gpurun_out/prof_all.ncu-rep
python tools/ncu_summary.py gpurun_out/prof_all.ncu-rep gpurun_out/ncu_all_kernels --engine-profile gpurun_out/profile_kernels.log --traffic-json gpurun_out/ncu_traffic.json > gpurun_out/ncu_summary.log 2>&1; echo "summary exit=$?"
sz=$(stat -c %s gpurun_out/prof_all.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 40000000 ]; then rm -f gpurun_out/prof_all.ncu-rep; echo "rep too large ($sz), removed after summarising"; fi
du -sh gpurun_out
