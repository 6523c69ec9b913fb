#!/bin/bash
# N-GPU validation (gpurun --gpus N): sharded parity worker, headline bench, sphere bench
N=${1:-8}; tag=${2:-r01}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 tests/shard_worker.py intel dlr manhattan100000 sphere2500 sphere200x200 > gpurun_out/shard_${tag}_n$N.log 2>&1; echo "shard worker rc=$?"
grep -E "shard ok|Error|error|assert" gpurun_out/shard_${tag}_n$N.log | cut -c1-300 | head -20
for wl in manhattan sphere; do
  out=gpurun_out/bench_${wl}_${tag}_n$N
  timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --workload $wl > $out.json 2> $out.err; echo "bench $wl rc=$?"
  python -c "
import json
d=json.loads([l for l in open('$out.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('n_gpus','ms_per_step','pcg_iterations_per_step','phase_ms','gpu_launches','chi2')}, d['roofline']['ms_per_launch'], d['e2e'])
"
done
