export PGO_COMM_TIMEOUT_S=3
for cfg in "2 100000" "3 100000" "4 100000" "3 300000"; do
  PGO_REPL_MAX_ROWS=700 timeout 120 python tools/debug_sharded_coarse.py $cfg 2>&1 | tail -1
done
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -x 2>&1 | tail -3
