#!/bin/bash
# Round-2 GPU call G: host-buffer pipeline after the hardware-queue fix (CUDA_DEVICE_MAX_CONNECTIONS=32, no idle second stream).
set -u
OUT=gpurun_out; mkdir -p $OUT
E="--steps 4 --warmup 2 --no-cpu --no-probe --verify-stride 0 --e2e-only --e2e-input packed"
for cfg in "3 1048576" "4 1048576" "6 1048576" "4 524288" "8 524288"; do set -- $cfg
  timeout 300 python bench.py $E --e2e-slots $1 --e2e-batch $2 > $OUT/g_e2e_s$1_b$2.json 2> $OUT/g_e2e_s$1_b$2.err; echo "e2e $cfg rc=$?"; python - $OUT/g_e2e_s$1_b$2.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print(d['input'], d['slots'], d['batch'], round(d['e2e_reads_per_s']/1e6,1))
PY
done
