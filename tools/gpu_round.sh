#!/bin/bash
# One GPU-box session: parity tests, bench, ncu launch list + full capture of the dominant kernel.
# usage (under gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$tag.log
tail -5 gpurun_out/pytest_gpu_$tag.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
cat gpurun_out/bench_$tag.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$tag.csv \
    python tools/profile_step.py --pcg-iters 16 > gpurun_out/ncu_launches_$tag.log 2>&1; echo "ncu list rc=$?"
bash tools/gpu_ncu_spmv.sh $tag
ls -la gpurun_out | tail -12
