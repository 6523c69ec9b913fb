#!/bin/bash
# r01k: fp32 storage of the preconditioner blocks + programmatic dependent launch: parity suites, then A/B of the headline step
tag=${1:-r01k}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_$tag.log | cut -c1-300
for pdl in 0 1; do
  echo "== PGO_PDL=$pdl  (configs: fp64 blocks ; fp32 blocks)"
  PGO_PDL=$pdl timeout 400 python tools/quick_perf.py --opts "amg_fp64_storage=1;amg_fp64_storage=0" 2>&1 | cut -c1-420 | tail -2
done
echo "== sphere PDL=1 fp64;fp32"; timeout 300 python tools/quick_perf.py --se3 --poses 250000 --opts "amg_fp64_storage=1;amg_fp64_storage=0" 2>&1 | cut -c1-420 | tail -2
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
cut -c1-1200 gpurun_out/bench_$tag.json
