#!/bin/bash
# r01n: TMA-staged sliced SpMV bring-up: ring-depth sweep, then the parity suites with it enabled
tag=${1:-r01n}
mkdir -p gpurun_out
timeout 300 python tools/tma_sweep.py --cfgs "0:0,2:3,3:4,4:6,6:8" 2>&1 | grep -v Warning | cut -c1-300 | tee gpurun_out/tma_sweep_$tag.log
timeout 300 python tools/tma_sweep.py --se3 --poses 250000 --cfgs "0:0,2:2,2:4,3:6" 2>&1 | grep -v Warning | cut -c1-300 | tee -a gpurun_out/tma_sweep_$tag.log
PGO_SPMV_TMA64=3 PGO_SPMV_TMA32=4 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_se3.py -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_$tag.log | cut -c1-300
