#!/bin/bash
# One gpurun call: the GPU test suite, the batch-1 latency check, one-batch class profile, bench.py (N=1).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"
tail -12 gpurun_out/pytest_gpu.log
timeout 300 python tools/profile_stream.py graph 2>&1 | tail -3
timeout 600 python tools/profile_step.py 1024 1 > gpurun_out/profile_step.log 2>&1; echo "profile_step exit=$?"; cat gpurun_out/profile_step.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit=$?"
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_n1.json"))
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], "roofline", {k:d["roofline"][k] for k in ("achieved","frac","share_of_step","other_classes_ms_per_step")}, "clocks", d["clocks"], "cpu", d.get("cpu_baseline",{}).get("value"), "stream", d.get("streaming"))
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/bench_n1.err
