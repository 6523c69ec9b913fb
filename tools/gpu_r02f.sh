#!/bin/bash
timeout 500 python tools/quick_perf.py --opts "sort_window=512;sort_window=1024;sort_window=4096;sort_window=8192;amg_aggregate_size=12;amg_aggregate_size=20;amg_aggregate_size=24;amg_aggregate_size=32;amg_aggregate_size=24,sort_window=4096" 2>&1 | grep -v Warning | cut -c1-330
