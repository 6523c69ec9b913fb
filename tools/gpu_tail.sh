#!/bin/bash
# coarse-tail kernel bring-up: parity suites, then the headline step with / without the tail and with 1..3 CTAs per SM
tag=${1:-tail}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_se3.py -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_$tag.log | cut -c1-300
for cfg in "PGO_TAIL=0" "PGO_TAIL_CTAS_PER_SM=1" "PGO_TAIL_CTAS_PER_SM=2" "PGO_TAIL_CTAS_PER_SM=3"; do
  echo "== $cfg"
  env $cfg timeout 300 python tools/quick_perf.py 2>&1 | cut -c1-400 | tail -2
done
echo "== sphere tail"; timeout 300 python tools/quick_perf.py --se3 --poses 250000 2>&1 | cut -c1-400 | tail -1
echo "== sphere no tail"; PGO_TAIL=0 timeout 300 python tools/quick_perf.py --se3 --poses 250000 2>&1 | cut -c1-400 | tail -1
