#!/bin/bash
# 1-GPU measurement session on the B200 box:  gpurun --timeout 2400 -- 'bash tools/gpu_session.sh TAG [parts]'
# parts (default: tests sweep bench traffic sanitize); others: multi strong scale create sweep2 quick ncugj bundled refine abwhile setup sanitize2
tag=${1:-r03}; parts=${2:-"tests sweep bench traffic sanitize"}
mkdir -p gpurun_out
for part in $parts; do case $part in
tests)
  timeout 2400 python -m pytest tests -m gpu -q -rs --durations=8 > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu_$tag.log;;
sweep)
  timeout 900 python tools/rtol_sweep.py 1e-8 1e-9 1e-10 1e-11 1e-12 1e-13 > gpurun_out/rtol_sweep_$tag.log 2>&1; echo "sweep rc=$?"; cat gpurun_out/rtol_sweep_$tag.log;;
bench)
  timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cut -c1-2500 gpurun_out/bench_$tag.json;;
traffic)
  # PGO_WHILE=0: ncu does not see the kernels inside the body of a conditional (WHILE) graph node; the chunked graph runs the same kernels
  PGO_WHILE=0 timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none \
      --profile-from-start off --csv --log-file gpurun_out/step_traffic_$tag.csv python tools/step_traffic.py > gpurun_out/step_traffic_$tag.log 2>&1; echo "traffic rc=$?"
  tail -2 gpurun_out/step_traffic_$tag.log; python tools/summarize_traffic.py gpurun_out/step_traffic_$tag.csv gpurun_out/step_traffic_$tag.json > gpurun_out/step_traffic_$tag.md; head -30 gpurun_out/step_traffic_$tag.md;;
multi)   # on a box with N >= 2 GPUs:  gpurun --gpus N -- 'bash tools/gpu_session.sh TAG multi'
  N=$(nvidia-smi -L | wc -l)
  timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_sharded.py -m gpu -q -rs > gpurun_out/pytest_multi_n${N}_$tag.log 2>&1; echo "pytest multi rc=$?"; tail -4 gpurun_out/pytest_multi_n${N}_$tag.log
  timeout 600 python bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_sp_n${N}_$tag.json 2> gpurun_out/bench_sp_n${N}_$tag.err; echo "bench single-process N=$N rc=$?"; cut -c1-1800 gpurun_out/bench_sp_n${N}_$tag.json
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 \
      > gpurun_out/bench_tr_n${N}_$tag.json 2> gpurun_out/bench_tr_n${N}_$tag.err; echo "bench torchrun N=$N rc=$?"; cut -c1-1800 gpurun_out/bench_tr_n${N}_$tag.json;;
strong)  # strong scaling of the 1M-pose graph on all GPUs of the box, both launch modes
  N=$(nvidia-smi -L | wc -l)
  timeout ${T1:-300} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 \
      > gpurun_out/bench_tr_n${N}_$tag.json 2> gpurun_out/bench_tr_n${N}_$tag.err; echo "bench torchrun N=$N rc=$?"
  timeout ${T1:-300} python bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_sp_n${N}_$tag.json 2> gpurun_out/bench_sp_n${N}_$tag.err; echo "bench single-process N=$N rc=$?"
  python - <<PY
import json
for f in ("bench_tr_n${N}_$tag", "bench_sp_n${N}_$tag"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "ms/step %.2f" % d["ms_per_step"], "its", d["pcg_iterations_per_step"], "phases", {k: round(v, 2) for k, v in d["phase_ms"].items()}, "e2e %.2f" % d["e2e"]["ms_per_step"], "parity", {k: v for k, v in (d.get("parity") or {}).items() if "err" in k})
    except Exception as e:
        print(f, "unreadable:", e)
PY
  ;;
scale)   # strong scaling (1M poses) as the driver runs it (torchrun) + weak scaling (1M poses per GPU) through the single-process handle
  N=$(nvidia-smi -L | wc -l)
  timeout ${T1:-600} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 \
      > gpurun_out/bench_tr_n${N}_$tag.json 2> gpurun_out/bench_tr_n${N}_$tag.err; echo "bench torchrun N=$N rc=$?"; cut -c1-300 gpurun_out/bench_tr_n${N}_$tag.json
  timeout ${T2:-900} python bench.py --gpus $N --poses $((N * 1000000)) --steps 3 --warmup 2 --no-secondary --no-cpu-baseline \
      > gpurun_out/bench_weak_n${N}_$tag.json 2> gpurun_out/bench_weak_n${N}_$tag.err; echo "bench weak N=$N rc=$?"; cut -c1-300 gpurun_out/bench_weak_n${N}_$tag.json
  python - <<PY
import json
for f in ("bench_tr_n${N}_$tag", "bench_weak_n${N}_$tag"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "poses", d["config"]["poses"], "ms/step %.2f" % d["ms_per_step"], "its", d["pcg_iterations_per_step"], "phases", {k: round(v, 2) for k, v in d["phase_ms"].items()}, "create_s %.1f" % d["create_s"], "parity", d.get("parity"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
  ;;
create)
  nproc | sed 's/^/host cores: /' | tee gpurun_out/create_$tag.log
  PGO_SYM_TIMING=1 timeout 300 python tools/time_create.py --repeats 3 2>&1 | tail -42 | tee -a gpurun_out/create_$tag.log
  PGO_HOST_THREADS=1 PGO_SYM_TIMING=1 timeout 300 python tools/time_create.py --repeats 2 2>&1 | tail -42 | sed 's/^/1 thread: /' | tee -a gpurun_out/create_$tag.log;;
sweep2)   # solver error (against this build's own refined solution) and error against the golden, around the bench tolerance
  timeout 900 python tools/rtol_sweep.py 2e-9 1e-9 5e-10 2e-10 1e-10 3e-11 --own > gpurun_out/rtol_sweep2_$tag.log 2>&1; echo "sweep2 rc=$?"; cut -c1-420 gpurun_out/rtol_sweep2_$tag.log;;
quick)
  timeout 300 python tools/quick_perf.py --opts pcg_rtol=1e-9 2>&1 | tail -1 | tee gpurun_out/quick_$tag.log
  timeout 300 python tools/quick_perf.py --se3 --poses 250000 --opts pcg_rtol=1e-9 2>&1 | tail -1 | sed "s/^/SE3 /" | tee -a gpurun_out/quick_$tag.log;;
ncugj)    # one source-level capture of the dense coarsest inversion
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_dense_invert -c 1 -f -o gpurun_out/gj_$tag python tools/quick_perf.py --poses 100000 > gpurun_out/ncugj_$tag.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncugj_$tag.log;;
bundled)  # the reference's own datasets (BASELINE configs[0..2]) + configs[4]
  : > gpurun_out/bench_bundled_$tag.jsonl
  for w in pose-pose pose-landmark intel dlr m3500 sphere2500 garage; do
    timeout 300 python bench.py --workload $w --steps 10 --warmup 3 >> gpurun_out/bench_bundled_$tag.jsonl 2>> gpurun_out/bench_bundled_$tag.err; echo "bench $w rc=$?"
  done
  python - <<PY
import json
for l in open("gpurun_out/bench_bundled_$tag.jsonl"):
    d = json.loads(l); c = d["cpu_baseline"] or {}
    print(d["config"]["workload"][:60], "| ms/step %.3f | pcg its %.0f | cpu s/GN it %s" % (d["ms_per_step"], d["pcg_iterations_per_step"], c.get("s_per_gn_iteration")))
PY
  ;;
refine)
  timeout 300 python tools/rtol_sweep.py 1e-9 1e-8 --opts=refine=1 2>&1 | tee gpurun_out/refine_$tag.log
  timeout 300 python tools/rtol_sweep.py 1e-9 --opts=refine=1,refine_rtol=1e-3 2>&1 | tail -1 | tee -a gpurun_out/refine_$tag.log
  timeout 300 python tools/rtol_sweep.py 1e-9 --opts=refine=1,refine_rtol=1e-5 2>&1 | tail -1 | tee -a gpurun_out/refine_$tag.log
  timeout 900 python -m pytest tests -m gpu -q -x -k "refine or variants or multi_gpu_handle_assembles" 2>&1 | tail -3 | tee -a gpurun_out/refine_$tag.log;;
abwhile)
  for v in 0 1; do PGO_WHILE=$v timeout 300 python tools/quick_perf.py --opts pcg_rtol=1e-9 2>&1 | tail -1 | sed "s/^/WHILE=$v /"; done | tee gpurun_out/while_$tag.log
  for v in 0 1; do PGO_WHILE=$v timeout 300 python tools/quick_perf.py --poses 100000 --opts pcg_rtol=1e-9 2>&1 | tail -1 | sed "s/^/100k WHILE=$v /"; done | tee -a gpurun_out/while_$tag.log;;
setup)   # what the hierarchy set-up of one GN step consists of (launch list of everything that is not the PCG loop)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv \
      -k regex:'galerkin|dense|to_float|coarse_pos|lever|invert_diag|assemble|build_hz|chi2|retract' \
      --log-file gpurun_out/setup_launches_$tag.csv python tools/step_traffic.py > gpurun_out/setup_launches_$tag.log 2>&1; echo "setup rc=$?"
  python tools/summarize_launches.py gpurun_out/setup_launches_$tag.csv | tee gpurun_out/setup_launches_$tag.md | head -40;;
sanitizeq)   # quick: memcheck + racecheck on the AMG path of two synthetic graphs (after a change of the set-up kernels)
  for tool in memcheck racecheck; do
    PGO_WHILE=0 timeout 200 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize.py 1 quick > gpurun_out/sanitize_${tool}_quick_$tag.log 2>&1; echo "$tool quick rc=$?"
    grep -E "sanitize |ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" gpurun_out/sanitize_${tool}_quick_$tag.log | head -8
  done;;
sanitize2)
  for tool in ${SAN_TOOLS:-memcheck racecheck}; do
    timeout 200 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize.py 2 > gpurun_out/sanitize_${tool}_n2_$tag.log 2>&1; echo "$tool n=2 rc=$?"
    grep -E "sanitize |ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" gpurun_out/sanitize_${tool}_n2_$tag.log | head -14
  done;;
sanitize)
  # PGO_WHILE=0: under compute-sanitizer a launch of the WHILE-node graph (device-side cudaGraphSetConditional) ends in error 700 with no
  # kernel-level record; the chunked graph runs the same kernels.  n = 1 only: the sanitizer serialises the kernels of a device, so
  # shards sharing a GPU time out by construction
  for tool in memcheck racecheck; do for n in 1; do
    PGO_WHILE=0 timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize.py $n > gpurun_out/sanitize_${tool}_n${n}_$tag.log 2>&1; echo "$tool n=$n rc=$?"
    grep -E "sanitize |ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" gpurun_out/sanitize_${tool}_n${n}_$tag.log | head -14
  done; done;;
esac; done
ls -la gpurun_out | tail -12
