#!/bin/bash
# r01t: full GPU suite, headline bench, sphere bench, the reference's bundled graphs
tag=${1:-r01t}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_$tag.log | cut -c1-300
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
cut -c1-700 gpurun_out/bench_$tag.json
timeout 900 python bench.py --workload sphere --steps 3 --warmup 3 > gpurun_out/bench_sphere_$tag.json 2> gpurun_out/bench_sphere_$tag.err; echo "bench sphere rc=$?"
cut -c1-400 gpurun_out/bench_sphere_$tag.json
for w in pose-pose pose-landmark intel dlr m3500 sphere2500 garage; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 >> gpurun_out/bench_bundled_$tag.jsonl 2>> gpurun_out/bench_bundled_$tag.err; echo "bench $w rc=$?"
done
python - <<'PY'
import json
for l in open("gpurun_out/bench_bundled_TAG.jsonl".replace("TAG", "'$tag'".strip("'"))):
    d = json.loads(l)
    print(d["config"]["workload"][:60], "| ms/step", round(d["ms_per_step"], 3), "| pcg", d["pcg_iterations_per_step"], "| edges/s", int(d["value"]), "| e2e ms", round(d["e2e"]["ms_per_step"], 3), "| cpu s/it", round(d["cpu_baseline"]["s_per_gn_iteration"], 4))
PY
