export PGO_COMM_TIMEOUT_S=3
for cfg in "2 100000" "3 100000" "2 100000 amg_kcycle3=0" "3 100000 amg_kcycle3=0"; do
  PGO_REPL_MAX_ROWS=700 timeout 120 python tools/debug_sharded_coarse.py $cfg 2>&1 | tail -1
done
PGO_REPL_MAX_ROWS=700 PGO_WHILE=0 timeout 120 python tools/debug_sharded_coarse.py 3 100000 2>&1 | tail -1
PGO_REPL_MAX_ROWS=700 PGO_PDL=0 timeout 120 python tools/debug_sharded_coarse.py 3 100000 2>&1 | tail -1
PGO_REPL_MAX_ROWS=700 PGO_WHILE=0 PGO_PDL=0 timeout 120 python tools/debug_sharded_coarse.py 3 100000 2>&1 | tail -1
PGO_REPL_MAX_ROWS=700 PGO_GJ_OLD=1 timeout 120 python tools/debug_sharded_coarse.py 3 100000 2>&1 | tail -1
timeout 120 python tools/debug_sharded_coarse.py 3 100000 2>&1 | tail -1
