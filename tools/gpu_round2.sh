#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit=$?"
tail -c 3000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit=$?"
timeout 300 python -m pytest tests/test_gpu_parity.py -q -s -m gpu -p no:cacheprovider 2>&1 | grep -E "parity|passed|failed" > gpurun_out/parity_report.txt
cat gpurun_out/parity_report.txt
timeout 300 python tools/profile_step.py > gpurun_out/profile_step.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_list.log 2>&1
cat gpurun_out/profile_step.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 39 -c 4 -o gpurun_out/prof_gemm python tools/profile_step.py > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit=$?"; tail -3 gpurun_out/ncu_full.log
