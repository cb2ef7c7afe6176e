#!/bin/bash
# Round-2 GPU call I: host-buffer pipeline with ordered batches (cfg.batch_order).
set -u
OUT=gpurun_out; mkdir -p $OUT
E="--steps 4 --warmup 2 --no-cpu --no-probe --verify-stride 0 --e2e-only --e2e-input packed"
for cfg in "3 1048576 1" "4 1048576 1" "3 2097152 1" "4 2097152 1" "3 2500000 1" "4 1048576 0"; do set -- $cfg
  timeout 300 python bench.py $E --e2e-slots $1 --e2e-batch $2 --batch-order $3 > $OUT/i_e2e_s$1_b$2_o$3.json 2> $OUT/i_e2e_s$1_b$2_o$3.err; echo "e2e $cfg rc=$?"; python - $OUT/i_e2e_s$1_b$2_o$3.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print(d['input'], d['slots'], d['batch'], round(d['e2e_reads_per_s']/1e6,1))
PY
done
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_chain.py -m gpu -q -x > $OUT/i_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/i_pytest.log; tail -4 $OUT/i_pytest.log
