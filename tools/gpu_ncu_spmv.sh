#!/bin/bash
# ncu --set full capture of the dominant kernel: the fine-level PCG SpMV k_spmv<3,0,1> on the 1M-pose graph
tag=${1:-r01}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k k_spmv -s 4 -c 3 -f \
    -o gpurun_out/prof_spmv_$tag python tools/profile_step.py --pcg-iters 16 > gpurun_out/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/ncu_full_$tag.log
