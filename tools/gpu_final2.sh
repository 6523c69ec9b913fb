#!/bin/bash
# final 1-GPU check of the round: full GPU suite, smoke, headline bench
tag=${1:-r02h}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$tag.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
cut -c1-330 gpurun_out/bench_$tag.json
timeout 600 python bench.py --workload sphere --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sphere_$tag.json 2> gpurun_out/bench_sphere_$tag.err; echo "bench sphere rc=$?"
cut -c1-330 gpurun_out/bench_sphere_$tag.json
