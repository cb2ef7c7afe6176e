#!/bin/bash
# Round-2 GPU call J: continuous pipeline across read sets; multi tests; default bench.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_chain.py -m gpu -q -x > $OUT/j_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/j_pytest.log; tail -4 $OUT/j_pytest.log
E="--steps 6 --warmup 2 --no-cpu --no-probe --verify-stride 0 --e2e-only --e2e-input packed"
for cfg in "3 1048576" "3 2097152" "4 2097152"; do set -- $cfg
  timeout 300 python bench.py $E --e2e-slots $1 --e2e-batch $2 > $OUT/j_e2e_s$1_b$2.json 2> $OUT/j_e2e_s$1_b$2.err; echo "e2e $cfg rc=$?"; python - $OUT/j_e2e_s$1_b$2.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print(d['input'], d['slots'], d['batch'], round(d['e2e_reads_per_s']/1e6,1))
PY
done
