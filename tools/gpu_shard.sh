#!/bin/bash
# Multi-GPU session (gpurun --gpus N): sharded parity worker + sharded bench at N ranks.
# usage: bash tools/gpu_shard.sh <N> <tag> [cases...]
N=${1:-2}; tag=${2:-r01}; shift 2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 tests/shard_worker.py "$@" > gpurun_out/shard_${tag}_n$N.log 2>&1; echo "shard worker rc=$?"
grep -E "shard ok|Error|error|assert" gpurun_out/shard_${tag}_n$N.log | head -20
timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${tag}_n$N.json 2> gpurun_out/bench_${tag}_n$N.err; echo "bench rc=$?"
cat gpurun_out/bench_${tag}_n$N.json; tail -5 gpurun_out/bench_${tag}_n$N.err
