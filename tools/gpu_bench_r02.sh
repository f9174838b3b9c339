#!/bin/bash
# One gpurun call (1 GPU): bench.py N=1, the reference arm, the ncu launch list of the bench command, and the ncu --set full
# capture of the block GEMMs at the bench shape -> per-GEMM DRAM traffic.  Outputs under gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
TAG=${1:-v1}
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_n1_$TAG.json 2> gpurun_out/bench_n1.err; echo "bench exit=$?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_bench_n1_$TAG.json"))
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], "e2e_cli", d.get("e2e_cli"))
    print("roofline", {k:d["roofline"][k] for k in ("achieved","frac","share_of_step","other_classes_ms_per_step","whole_step_frac")})
    print("clocks", d["clocks"], "cpu", d.get("cpu_baseline",{}).get("value"))
    print("stereo", d.get("stereo")); print("one_shot", d.get("one_shot_10s")); print("stream", json.dumps(d.get("streaming"))[:900])
except Exception as e:
    print("bench parse failed", e)
PY
tail -5 gpurun_out/bench_n1.err
if [ "$2" = "ncu" ]; then
  timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:gemm2_bf16 -c 24 -o gpurun_out/prof_gemm python tools/profile_step.py 1024 1 > gpurun_out/ncu_gemm.log 2>&1
  echo "ncu gemm exit=$?"; tail -2 gpurun_out/ncu_gemm.log
  python tools/ncu_summary.py gpurun_out/prof_gemm.ncu-rep gpurun_out/r02_ncu_block_gemms_bench_shape > /dev/null 2>&1; echo "summary exit=$?"
  python tools/ncu_gemm_traffic.py gpurun_out/r02_ncu_block_gemms_bench_shape.json gpurun_out/ncu_traffic.json
  rm -f gpurun_out/prof_gemm.ncu-rep
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_ncu_launches_bench_cmd.csv python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list exit=$?"
  python tools/launch_shares.py gpurun_out/r02_ncu_launches_bench_cmd.csv > gpurun_out/r02_ncu_launch_shares_bench_cmd.txt 2>&1; cat gpurun_out/r02_ncu_launch_shares_bench_cmd.txt
  gzip -f gpurun_out/r02_ncu_launches_bench_cmd.csv
fi
du -sh gpurun_out
