timeout 600 python tools/quick_perf.py --poses 4000000 --opts pcg_rtol=1e-9 2>&1 | tail -1
timeout 600 python tools/quick_perf.py --poses 2000000 --opts pcg_rtol=1e-9 2>&1 | tail -1
