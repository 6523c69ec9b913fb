for v in 1.21 1.5 1.7; do PGO_OMEGA_RHO=$v timeout 300 python tools/quick_perf.py --opts pcg_rtol=1e-9 2>&1 | tail -1 | sed "s/^/OMEGA_RHO=$v /"; done
for v in 1.21 1.5; do PGO_OMEGA_RHO=$v timeout 300 python tools/quick_perf.py --se3 --poses 250000 --opts pcg_rtol=1e-9 2>&1 | tail -1 | sed "s/^/SE3 OMEGA_RHO=$v /"; done
PGO_OMEGA_RHO=1.5 timeout 300 python tools/rtol_sweep.py 1e-9 1e-10 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
