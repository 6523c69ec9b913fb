#!/bin/bash
# final 1-GPU session of a round: full GPU suite, smoke, headline bench, sphere bench, launch lists (cold + warm caches)
tag=${1:-r02c}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_$tag.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/bench_$tag.json
timeout 900 python bench.py --workload sphere --steps 5 --warmup 3 > gpurun_out/bench_sphere_$tag.json 2> gpurun_out/bench_sphere_$tag.err; echo "bench sphere rc=$?"
cut -c1-300 gpurun_out/bench_sphere_$tag.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$tag.csv \
    python tools/profile_step.py --pcg-iters 12 > gpurun_out/ncu_launches_$tag.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 4000 --csv --log-file gpurun_out/launches_warm_$tag.csv \
    python tools/profile_step.py --pcg-iters 12 > gpurun_out/ncu_launches_warm_$tag.log 2>&1; echo "ncu warm list rc=$?"
