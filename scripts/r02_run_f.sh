#!/bin/bash
# Round-2 GPU call F: chain tests after the chain-overflow fix, host-buffer pipeline timelines (cs_multi_trace).
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_chain.py -m gpu -q > $OUT/f_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/f_pytest.log; tail -4 $OUT/f_pytest.log
E="--steps 4 --warmup 2 --no-cpu --no-probe --verify-stride 0 --e2e-only --e2e-input bytes"
for cfg in "3 1048576" "6 1048576" "4 524288"; do set -- $cfg
  timeout 300 python bench.py $E --e2e-slots $1 --e2e-batch $2 > $OUT/f_e2e_s$1_b$2.json 2> $OUT/f_e2e_s$1_b$2.err; echo "e2e $cfg rc=$?"; done
