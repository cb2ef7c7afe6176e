#!/bin/bash
# Round-2 GPU call C: tests again, host-buffer pipeline experiments, literal-kernel build variants on cfg4 / cfg2.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/c_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/c_pytest.log; tail -4 $OUT/c_pytest.log
E="--steps 6 --warmup 2 --no-cpu --no-probe --verify-stride 0 --e2e-only"
for cfg in "3 1048576" "4 1048576" "6 1048576" "4 524288" "3 2097152"; do set -- $cfg
  timeout 300 python bench.py $E --e2e-slots $1 --e2e-batch $2 > $OUT/c_e2e_s$1_b$2.json 2> $OUT/c_e2e_s$1_b$2.err; echo "e2e $cfg rc=$?"; done
for tag in "" e16b2 e6b4 e4b4 q4 q8; do
  COMPSEED_LIB_TAG=$tag timeout 300 python scripts/r02_cfg4.py > $OUT/c_cfg4_${tag:-base}.json 2> $OUT/c_cfg4_${tag:-base}.err; echo "cfg4 $tag rc=$?"; done
V="--reads 4000000 --steps 3 --warmup 2 --no-cpu --no-e2e --no-probe --verify-stride 0"
for tag in "" e16b2 e6b4 q4; do
  COMPSEED_LIB_TAG=$tag timeout 200 python bench.py $V > $OUT/c_var_${tag:-base}.json 2> $OUT/c_var_${tag:-base}.err; echo "var $tag rc=$?"; done
ls $OUT | grep "^c_" | head -40
