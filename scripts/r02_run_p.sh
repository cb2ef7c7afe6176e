#!/bin/bash
# Round-2 GPU call P: calls with a long non-unique forward match go to k_seed_walk (CS_WALK_L) -- parity tests, cfg2 4 M-read step with
# and without it, event counters, cfg4.
set -u
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q ) > $OUT/p_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/p_pytest.log; tail -6 $OUT/p_pytest.log
S="--reads 4000000 --steps 3 --warmup 2 --no-cpu --no-e2e --no-probe --verify-stride 0"
for tag in "" nowl; do
  COMPSEED_LIB_TAG=$tag timeout 300 python bench.py $S > $OUT/p_small_$tag.json 2> $OUT/p_small_$tag.err; echo "small '$tag' rc=$?"
  COMPSEED_LIB_TAG=$tag timeout 300 python scripts/r02_cfg4.py > $OUT/p_cfg4_$tag.json 2> $OUT/p_cfg4_$tag.err; echo "cfg4 '$tag' rc=$?"; cat $OUT/p_cfg4_$tag.json
done
COMPSEED_LIB_TAG=stats timeout 300 python scripts/spec_stats.py 1000000 > $OUT/p_stats.txt 2>&1; echo "stats rc=$?"; head -8 $OUT/p_stats.txt
