#!/bin/bash
# Round-2 GPU call N: repeat-length array + diagonal speculation -- parity tests, full-size tests (index verify incl. rep), default bench
# line, variants (3 CTAs per SM without spills; speculation off), event counters of the stats build, cfg4.
set -u
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_multi.py -m gpu -x -q ) > $OUT/n_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/n_pytest.log; tail -6 $OUT/n_pytest.log
( time timeout 600 python bench.py ) > $OUT/n_bench.json 2> $OUT/n_bench.err; echo "bench rc=$?"; tail -3 $OUT/n_bench.err
S="--reads 4000000 --steps 3 --warmup 2 --no-cpu --no-e2e --no-probe --verify-stride 0"
for tag in "" mb3 nospec; do
  COMPSEED_LIB_TAG=$tag timeout 300 python bench.py $S > $OUT/n_small_$tag.json 2> $OUT/n_small_$tag.err; echo "small '$tag' rc=$?"
done
COMPSEED_LIB_TAG=stats timeout 300 python scripts/spec_stats.py 1000000 > $OUT/n_stats.txt 2>&1; echo "stats rc=$?"; head -60 $OUT/n_stats.txt
timeout 300 python scripts/r02_cfg4.py > $OUT/n_cfg4.json 2> $OUT/n_cfg4.err; echo "cfg4 rc=$?"; cat $OUT/n_cfg4.json
