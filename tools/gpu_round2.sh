#!/bin/bash
# 1-GPU session: full GPU test-suite, headline bench, sphere (SE3) bench, ncu --set full of the fine SpMV for both block sizes
tag=${1:-r01}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
cut -c1-900 gpurun_out/bench_$tag.json
timeout 900 python bench.py --workload sphere --steps 3 --warmup 3 > gpurun_out/bench_sphere_$tag.json 2> gpurun_out/bench_sphere_$tag.err; echo "bench sphere rc=$?"
cut -c1-1500 gpurun_out/bench_sphere_$tag.json
timeout 600 ncu --set full --clock-control none --import-source on -k k_spmv -s 4 -c 2 -f -o gpurun_out/prof_spmv3_$tag \
    python tools/profile_step.py --pcg-iters 8 > gpurun_out/ncu_full3_$tag.log 2>&1; echo "ncu3 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k k_spmv -s 4 -c 2 -f -o gpurun_out/prof_spmv6_$tag \
    python tools/profile_step.py --se3 --poses 250000 --pcg-iters 8 > gpurun_out/ncu_full6_$tag.log 2>&1; echo "ncu6 rc=$?"
ls -la gpurun_out | tail -8
