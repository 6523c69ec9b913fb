#!/bin/bash
# SE3 session: SE3 parity tests (all failures shown), then a perf probe on BASELINE configs[4] (250k-pose sphere)
tag=${1:-se3}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_se3.py -m gpu -q > gpurun_out/pytest_se3_$tag.log 2>&1; echo "se3 rc=$?"
tail -60 gpurun_out/pytest_se3_$tag.log | cut -c1-300
timeout 600 python tools/quick_perf.py --se3 --poses 250000 --opts "${2:-;preconditioner=0}" > gpurun_out/perf_se3_$tag.log 2>&1; echo "perf rc=$?"
cut -c1-700 gpurun_out/perf_se3_$tag.log | tail
