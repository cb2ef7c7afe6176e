#!/bin/bash
# Round-2 GPU call Q: extension kernel with its DP rows and a packed copy of the query in shared memory -- parity tests, timing against
# the HBM-scratch kernel, CTA sizes 32 / 64 / 128.
set -u
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_gpu_bsw.py -m gpu -x -q ) > $OUT/q_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/q_pytest.log; tail -6 $OUT/q_pytest.log
for tag in "" b32 b128; do COMPSEED_LIB_TAG=$tag timeout 300 python scripts/bsw_time.py 2000000 8 >> $OUT/q_bsw_time.log 2>&1; done; cat $OUT/q_bsw_time.log
