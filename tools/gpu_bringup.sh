#!/bin/bash
# One gpurun call: stage-by-stage diagnostics (separate processes, each under its own timeout), then pytest -m gpu.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for st in gemm attn vq e2e; do
  echo "===== diag $st =====" >> gpurun_out/diag.log
  timeout 240 python tools/diag.py $st >> gpurun_out/diag.log 2>&1
  echo "exit=$?" >> gpurun_out/diag.log
done
timeout 1200 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
cat gpurun_out/diag.log | tail -80
