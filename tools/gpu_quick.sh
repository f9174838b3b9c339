#!/bin/bash
# quick validation: gemm diag, full gpu test-suite, one profiled batch
mkdir -p gpurun_out
timeout 300 python tools/diag.py gemm > gpurun_out/diag_gemm.log 2>&1; echo "diag exit=$?"; cat gpurun_out/diag_gemm.log
timeout 900 python -m pytest tests -q -m gpu --tb=short --durations=6 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tools/profile_step.py > gpurun_out/profile_step.log 2>&1; cat gpurun_out/profile_step.log
timeout 300 python tools/profile_step.py > gpurun_out/profile_step_b.log 2>&1; tail -2 gpurun_out/profile_step_b.log
