#!/bin/bash
for pdl in 1 0; do echo "== PGO_PDL=$pdl"; PGO_PDL=$pdl timeout 300 python tools/spmv_sweep.py 2>&1 | grep -v Warning | cut -c1-300 | tail -2; done
