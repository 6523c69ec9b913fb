#!/bin/bash
# r01l: sliced SpMV variants (entries per iteration, register cap), fp64 and fp32 block storage; coarse-solve timing per level
tag=${1:-r01l}
mkdir -p gpurun_out
timeout 600 python tools/spmv_sweep.py --us "" --steps "1:1:1,1:1:12" 2>&1 | grep -v Warning | cut -c1-400 | tee gpurun_out/spmv_sweep_$tag.log
timeout 600 python tools/spmv_sweep.py --se3 --poses 250000 --us "" --minbs "" --steps "1:1:1" 2>&1 | grep -v Warning | cut -c1-400 | tee -a gpurun_out/spmv_sweep_$tag.log
