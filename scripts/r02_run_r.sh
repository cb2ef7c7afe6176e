#!/bin/bash
# Round-2 GPU call R: where the literal kernel's tasks come from (stats build), cfg2 and the repeat-rich reference.
set -u
OUT=gpurun_out; mkdir -p $OUT
COMPSEED_LIB_TAG=stats timeout 300 python scripts/spec_stats.py 1000000 > $OUT/r_stats_cfg2.txt 2>&1; echo "stats cfg2 rc=$?"; sed -n 1,16p $OUT/r_stats_cfg2.txt
COMPSEED_LIB_TAG=stats timeout 300 python scripts/spec_stats.py 400000 20000000 repeat > $OUT/r_stats_cfg4.txt 2>&1; echo "stats cfg4 rc=$?"; sed -n 1,45p $OUT/r_stats_cfg4.txt
