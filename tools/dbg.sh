timeout 900 python tools/quick_perf.py --poses 2000000 --opts "pcg_rtol=1e-9,amg_kcycle3=2;pcg_rtol=1e-9,amg_dense_max=1024;pcg_rtol=1e-9,amg_kcycle3=2,amg_dense_max=1024;pcg_rtol=1e-9,amg_kcycle3=3" 2>&1 | grep cfg
timeout 900 python tools/quick_perf.py --poses 4000000 --opts "pcg_rtol=1e-9,amg_kcycle3=2;pcg_rtol=1e-9,amg_kcycle3=3;pcg_rtol=1e-9,amg_kcycle3=2,amg_aggregate_size=24" 2>&1 | grep cfg
timeout 900 python tools/quick_perf.py --poses 1000000 --opts "pcg_rtol=1e-9,amg_kcycle3=2" 2>&1 | grep cfg
