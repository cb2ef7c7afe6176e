#!/bin/bash
# Round-2 GPU call A: GPU tests, the default bench line (parity + in-run probe), the reference arm, and ncu evidence
# (launch list + full capture of the seeding kernels) on a reduced step.  Run under gpurun from the repo root.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/a_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/a_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/a_pytest.log
tail -5 $OUT/a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/a_bench.json 2> $OUT/a_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/a_bench_ref.json 2> $OUT/a_bench_ref.err; echo "ref rc=$?"
SMALL="python bench.py --reads 2000000 --steps 1 --warmup 1 --no-cpu --no-e2e --no-probe --verify-stride 0"
timeout 300 $SMALL > $OUT/a_small.json 2> $OUT/a_small.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/a_launches.csv $SMALL > $OUT/a_ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 300 $SMALL > $OUT/a_small2.json 2> $OUT/a_small2.err &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_seed -s 4 -c 4 -o $OUT/a_prof_seed -f $SMALL > $OUT/a_ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la $OUT | tail -20
