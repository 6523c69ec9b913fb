#!/bin/bash
# r01q: ncu launch list of one GN step (PCG capped) + --set full captures of the sliced SpMV (fp64 + fp32 blocks) and the dense inverse
tag=${1:-r01q}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$tag.csv \
    python tools/profile_step.py --pcg-iters 12 > gpurun_out/ncu_launches_$tag.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv -s 6 -c 4 -f -o gpurun_out/prof_spmv_$tag \
    python tools/profile_step.py --pcg-iters 8 > gpurun_out/ncu_full_$tag.log 2>&1; echo "ncu spmv rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense_invert -c 1 -f -o gpurun_out/prof_inv_$tag \
    python tools/profile_step.py --pcg-iters 2 > gpurun_out/ncu_inv_$tag.log 2>&1; echo "ncu inv rc=$?"
ls -la gpurun_out | tail -6
