#!/bin/bash
tag=${1:-r01s}
mkdir -p gpurun_out
timeout 400 python tools/spmv_sweep.py --opts "amg_kcycle3=0;amg_kcycle3=1;amg_kcycle3=2" 2>&1 | grep -v Warning | cut -c1-400 | tail -6 | tee gpurun_out/step_$tag.log
timeout 400 python tools/spmv_sweep.py --se3 --poses 250000 --opts "amg_kcycle3=0;amg_kcycle3=1;amg_kcycle3=2" 2>&1 | grep -v Warning | cut -c1-400 | tail -6 | tee -a gpurun_out/step_$tag.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_se3.py -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_$tag.log | cut -c1-300
