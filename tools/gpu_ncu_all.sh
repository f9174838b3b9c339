#!/bin/bash
# ncu --set full over every kernel of one encode + one decode pass (tools/profile_kernels.py)
mkdir -p gpurun_out
timeout 300 python tools/profile_kernels.py > gpurun_out/profile_kernels.log 2>&1; echo "plain exit=$?"
tail -2 gpurun_out/profile_kernels.log
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_all \
  python tools/profile_kernels.py > gpurun_out/ncu_all.log 2>&1
echo "ncu exit=$?"; tail -2 gpurun_out/ncu_all.log; ls -la gpurun_out/prof_all.ncu-rep
