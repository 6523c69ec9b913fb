#!/bin/bash
tag=${1:-r02b}
mkdir -p gpurun_out
timeout 200 python tools/spmv_sweep.py 2>&1 | grep -v Warning | cut -c1-300 | tail -2 | tee -a gpurun_out/step_$tag.log
timeout 300 python tools/spmv_sweep.py --se3 --poses 250000 2>&1 | grep -v Warning | cut -c1-300 | tail -2 | tee -a gpurun_out/step_$tag.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_se3.py -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_$tag.log | cut -c1-300
