#!/bin/bash
# Multi-GPU session: sharded parity worker (optional) + sharded bench with 2- and 3-step K-cycle
# usage: bash tools/gpu_shard2.sh <N> <tag> <parity:0|1>
N=${1:-2}; tag=${2:-r01}; par=${3:-1}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$par" = "1" ]; then
  timeout 900 $TR --master-port 29511 tests/shard_worker.py > gpurun_out/shard_${tag}_n$N.log 2>&1; echo "shard worker rc=$?"
  grep -E "shard ok|Error|error|assert" gpurun_out/shard_${tag}_n$N.log | head -20 | cut -c1-220
fi
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${tag}_n$N.json 2> gpurun_out/bench_${tag}_n$N.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_${tag}_n$N.json')); print('default:', d['ms_per_step'], d['pcg_iterations_per_step'], d['phase_ms'], d['roofline']['ms_per_launch'])"
timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 --opts amg_kcycle3=0 > gpurun_out/bench_${tag}_k2_n$N.json 2> gpurun_out/bench_${tag}_k2_n$N.err; echo "bench k2 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_${tag}_k2_n$N.json')); print('kcycle3=0:', d['ms_per_step'], d['pcg_iterations_per_step'], d['phase_ms'], d['roofline']['ms_per_launch'])"
tail -3 gpurun_out/bench_${tag}_n$N.err | cut -c1-300
