#!/bin/bash
# Round-2 GPU call D: all GPU tests (chaining included), default bench, pipeline depth sweep, literal-kernel variants.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -q > $OUT/d_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/d_pytest.log; tail -4 $OUT/d_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/d_bench.json 2> $OUT/d_bench.err; echo "bench rc=$?"
E="--steps 6 --warmup 2 --no-cpu --no-probe --verify-stride 0 --e2e-only"
for cfg in "3 1048576" "4 1048576" "6 1048576" "8 524288"; do set -- $cfg
  timeout 300 python bench.py $E --e2e-slots $1 --e2e-batch $2 > $OUT/d_e2e_s$1_b$2.json 2> $OUT/d_e2e_s$1_b$2.err; echo "e2e $cfg rc=$?"; done
for tag in "" e16b2 e6b4 q4; do
  COMPSEED_LIB_TAG=$tag timeout 300 python scripts/r02_cfg4.py > $OUT/d_cfg4_${tag:-base}.json 2> $OUT/d_cfg4_${tag:-base}.err; echo "cfg4 $tag rc=$?"; done
V="--reads 4000000 --steps 3 --warmup 2 --no-cpu --no-e2e --no-probe --verify-stride 0"
for tag in e16b2 e6b4 q4; do
  COMPSEED_LIB_TAG=$tag timeout 200 python bench.py $V > $OUT/d_var_${tag:-base}.json 2> $OUT/d_var_${tag:-base}.err; echo "var $tag rc=$?"; done
