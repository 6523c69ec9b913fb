#!/bin/bash
# r01v: launch list with the caches left as the running step leaves them (--cache-control none): in-context kernel durations
tag=${1:-r01v}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 4000 --csv --log-file gpurun_out/launches_warm_$tag.csv \
    python tools/profile_step.py --pcg-iters 12 > gpurun_out/ncu_launches_warm_$tag.log 2>&1; echo "ncu list rc=$?"
tail -2 gpurun_out/ncu_launches_warm_$tag.log
