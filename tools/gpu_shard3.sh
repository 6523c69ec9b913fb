#!/bin/bash
# usage: bash tools/gpu_shard3.sh <N> <tag> [parity cases...]   -- sharded parity on the given cases + the sharded headline bench
N=${1:-8}; tag=${2:-r01}; shift 2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ $# -gt 0 ]; then
  timeout 600 $TR --master-port 29511 tests/shard_worker.py "$@" > gpurun_out/shard_${tag}_n$N.log 2>&1; echo "shard worker rc=$?"
  grep -E "shard ok|Error|error|assert" gpurun_out/shard_${tag}_n$N.log | head -20 | cut -c1-260
fi
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${tag}_n$N.json 2> gpurun_out/bench_${tag}_n$N.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/bench_${tag}_n$N.json
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_${tag}_n$N.json') if l.startswith('{')][0]); print('N=$N:', d['ms_per_step'], d['pcg_iterations_per_step'], d['phase_ms'], d['roofline']['ms_per_launch'], d['e2e']['ms_per_step'])"
