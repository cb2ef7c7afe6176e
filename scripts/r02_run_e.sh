#!/bin/bash
# Round-2 GPU call E: chain + SAM tests, batch-size dependence (device-resident and host-buffer), default bench.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_chain.py tests/test_gpu_sam_identity.py tests/test_gpu_multi.py -m gpu -q > $OUT/e_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/e_pytest.log; tail -4 $OUT/e_pytest.log
V="--steps 5 --warmup 3 --no-cpu --no-e2e --no-probe --verify-stride 0"
for n in 1000000 2000000; do timeout 200 python bench.py $V --reads $n > $OUT/e_dev_$n.json 2> $OUT/e_dev_$n.err; echo "dev $n rc=$?"; done
E="--steps 6 --warmup 2 --no-cpu --no-probe --verify-stride 0 --e2e-only --e2e-input bytes"
for cfg in "3 1048576" "3 2097152" "3 4194304" "6 1048576" "2 5000000"; do set -- $cfg
  timeout 300 python bench.py $E --e2e-slots $1 --e2e-batch $2 > $OUT/e_e2e_s$1_b$2.json 2> $OUT/e_e2e_s$1_b$2.err; echo "e2e $cfg rc=$?"; done
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/e_bench.json 2> $OUT/e_bench.err; echo "bench rc=$?"
