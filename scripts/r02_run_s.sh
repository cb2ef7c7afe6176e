#!/bin/bash
# Round-2 GPU call S: k_seed_walk with the K-mer probe after the longest entry of every list and the second-pass follow-ups inline --
# parity tests, cfg2 4 M-read step and cfg4 (default build; walk kernel at 3 CTAs per SM), where the literal tasks come from now.
set -u
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q ) > $OUT/s_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/s_pytest.log; tail -6 $OUT/s_pytest.log
S="--reads 4000000 --steps 3 --warmup 2 --no-cpu --no-e2e --no-probe --verify-stride 0"
for tag in "" wmb3; do
  COMPSEED_LIB_TAG=$tag timeout 300 python bench.py $S > $OUT/s_small_$tag.json 2> $OUT/s_small_$tag.err; echo "small '$tag' rc=$?"
  COMPSEED_LIB_TAG=$tag timeout 300 python scripts/r02_cfg4.py > $OUT/s_cfg4_$tag.json 2> $OUT/s_cfg4_$tag.err; echo "cfg4 '$tag' rc=$?"; cat $OUT/s_cfg4_$tag.json
done
COMPSEED_LIB_TAG=stats timeout 300 python scripts/spec_stats.py 1000000 > $OUT/s_stats_cfg2.txt 2>&1; echo "stats cfg2 rc=$?"; sed -n 1,16p $OUT/s_stats_cfg2.txt
