#!/bin/bash
# Scaling session (gpurun --gpus 8): the headline bench at N = 1, 2, 4, 8 back to back, as the driver does at round end.
# usage: bash tools/gpu_scale.sh <tag>
tag=${1:-scale}
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + N)) bench.py --gpus $N --steps 5 --warmup 3 \
      > gpurun_out/bench_${tag}_n$N.json 2> gpurun_out/bench_${tag}_n$N.err
done
python - <<PY
import json
base = None
for n in (1, 2, 4, 8):
    d = json.loads([l for l in open(f"gpurun_out/bench_${tag}_n{n}.json") if l.startswith("{")][0])
    base = base or d["value"]
    print(f"N={n}: {d['ms_per_step']:.2f} ms/step, {d['pcg_iterations_per_step']:.0f} PCG its, {d['value']/1e6:.1f} M edges/s, speed-up {d['value']/base:.2f}, e2e {d['e2e']['ms_per_step']:.2f} ms")
PY
