export PGO_COMM_TIMEOUT_S=10
PGO_REPL_MAX_ROWS=50000 timeout 300 python tools/debug_sharded_coarse.py 4 1000000 pcg_max_iterations=600 2>&1 | tail -2
timeout 300 python tools/debug_sharded_coarse.py 4 1000000 pcg_max_iterations=600 2>&1 | tail -2
PGO_REPL_MAX_ROWS=50000 timeout 300 python tools/debug_sharded_coarse.py 2 1000000 pcg_max_iterations=600 2>&1 | tail -2
PGO_REPL_MAX_ROWS=20000 timeout 300 python tools/debug_sharded_coarse.py 4 400000 pcg_max_iterations=600 2>&1 | tail -2
PGO_REPL_MAX_ROWS=20000 PGO_PDL=0 timeout 300 python tools/debug_sharded_coarse.py 4 400000 pcg_max_iterations=600 2>&1 | tail -2
