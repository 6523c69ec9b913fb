#!/bin/bash
for c in 1 2 3 4 8; do echo "== PGO_CHUNK=$c"; PGO_CHUNK=$c timeout 200 python tools/quick_perf.py 2>&1 | grep -v Warning | cut -c1-330 | tail -1; done
for c in 1 2 4; do echo "== sphere PGO_CHUNK=$c"; PGO_CHUNK=$c timeout 200 python tools/quick_perf.py --se3 --poses 250000 2>&1 | grep -v Warning | cut -c1-330 | tail -1; done
