#!/bin/bash
tag=${1:-r01x}
mkdir -p gpurun_out
for cfg in "PGO_DEEP=0" "PGO_DEEP_DENSE=1" "PGO_DEEP_DENSE=0"; do
  echo "== $cfg"; env $cfg timeout 200 python tools/spmv_sweep.py 2>&1 | grep -v Warning | cut -c1-300 | tail -2 | tee -a gpurun_out/step_$tag.log
done
for cfg in "PGO_DEEP=0" "PGO_DEEP_DENSE=1" "PGO_DEEP_DENSE=0"; do
  echo "== sphere $cfg"; env $cfg timeout 200 python tools/spmv_sweep.py --se3 --poses 250000 2>&1 | grep -v Warning | cut -c1-300 | tail -2 | tee -a gpurun_out/step_$tag.log
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_se3.py -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_$tag.log | cut -c1-300
