#!/bin/bash
# Round-2 GPU call B: all GPU tests, default bench (C pipeline e2e, two probes), kernel variants on 4M reads, event counters.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $OUT/b_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/b_pytest.log
tail -8 $OUT/b_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/b_bench.json 2> $OUT/b_bench.err; echo "bench rc=$?"
V="--reads 4000000 --steps 3 --warmup 2 --no-cpu --no-e2e --no-probe --verify-stride 0"
timeout 200 python bench.py $V > $OUT/b_var_default.json 2> $OUT/b_var_default.err; echo "default rc=$?"
COMPSEED_LIB_TAG=fw1 timeout 200 python bench.py $V > $OUT/b_var_fw1.json 2> $OUT/b_var_fw1.err; echo "fw1 rc=$?"
COMPSEED_LIB_TAG=fw2 timeout 200 python bench.py $V > $OUT/b_var_fw2.json 2> $OUT/b_var_fw2.err; echo "fw2 rc=$?"
timeout 200 python bench.py $V --isa-intv 2 > $OUT/b_var_isa2.json 2> $OUT/b_var_isa2.err; echo "isa2 rc=$?"
timeout 200 python bench.py $V --no-overlap > $OUT/b_var_noov.json 2> $OUT/b_var_noov.err; echo "noov rc=$?"
timeout 200 python bench.py $V --lit-ctas 1 > $OUT/b_var_lit1.json 2> $OUT/b_var_lit1.err; echo "lit1 rc=$?"
timeout 200 python bench.py $V --l2-persist-mb 64 > $OUT/b_var_l2p.json 2> $OUT/b_var_l2p.err; echo "l2p rc=$?"
COMPSEED_LIB_TAG=stats timeout 300 python scripts/spec_stats.py 4000000 > $OUT/b_stats_cfg2.log 2>&1; echo "stats rc=$?"
COMPSEED_LIB_TAG=stats timeout 300 python scripts/spec_stats.py 400000 20000000 repeat > $OUT/b_stats_cfg4.log 2>&1; echo "stats4 rc=$?"
ls -la $OUT | grep " b_"
