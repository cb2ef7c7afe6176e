#!/bin/bash
# Round-2 GPU call X: BASELINE.json configs 4-5 with the final build (scripts/config_sweep.py).
set -u
OUT=gpurun_out; mkdir -p $OUT
( time timeout 1500 python scripts/config_sweep.py ) > $OUT/x_config_sweep.json 2> $OUT/x_config_sweep.err; echo "sweep rc=$?"; tail -3 $OUT/x_config_sweep.err | cut -c1-300
