#!/bin/bash
# Multi-GPU bench only (gpurun --gpus N): bash tools/gpu_shard_bench.sh <N> <tag> [ENV=VAL ...]
N=${1:-2}; tag=${2:-r01}; shift 2
mkdir -p gpurun_out
for kv in "$@"; do export "$kv"; done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
out=gpurun_out/bench_${tag}_n$N
timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > $out.json 2> $out.err; echo "bench rc=$?"
python -c "
import json
d=json.loads([l for l in open('$out.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('n_gpus','ms_per_step','pcg_iterations_per_step','phase_ms','gpu_launches')}, d['roofline']['ms_per_launch'])
"
tail -3 $out.err | cut -c1-400
