#!/bin/bash
# Round-2 GPU call H: host-buffer pipeline with result copies on their own stream; all GPU tests.
set -u
OUT=gpurun_out; mkdir -p $OUT
E="--steps 4 --warmup 2 --no-cpu --no-probe --verify-stride 0 --e2e-only --e2e-input packed"
for cfg in "2 1048576" "3 1048576" "4 1048576" "3 524288" "4 524288" "6 262144" "3 2097152"; do set -- $cfg
  timeout 300 python bench.py $E --e2e-slots $1 --e2e-batch $2 > $OUT/h_e2e_s$1_b$2.json 2> $OUT/h_e2e_s$1_b$2.err; echo "e2e $cfg rc=$?"; python - $OUT/h_e2e_s$1_b$2.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print(d['input'], d['slots'], d['batch'], round(d['e2e_reads_per_s']/1e6,1))
PY
done
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/h_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/h_pytest.log; tail -4 $OUT/h_pytest.log
