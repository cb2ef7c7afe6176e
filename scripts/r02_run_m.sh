#!/bin/bash
# Round-2 GPU call M: the whole GPU test suite, the default bench line (as the driver runs it), the reference arm, cfg4 timing,
# and the ncu evidence of the current build: launch list of two reduced steps + full capture of the dominant kernel.
set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/m_gpu.txt 2>&1; nproc >> $OUT/m_gpu.txt
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $OUT/m_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/m_pytest.log; tail -6 $OUT/m_pytest.log
( time timeout 900 python bench.py ) > $OUT/m_bench.json 2> $OUT/m_bench.err; echo "bench rc=$?"; tail -3 $OUT/m_bench.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > $OUT/m_bench_ref.json 2> $OUT/m_bench_ref.err; echo "ref rc=$?"
timeout 300 python scripts/r02_cfg4.py > $OUT/m_cfg4.json 2> $OUT/m_cfg4.err; echo "cfg4 rc=$?"; cat $OUT/m_cfg4.json
SMALL="python bench.py --reads 2000000 --steps 2 --warmup 3 --no-cpu --no-e2e --no-probe --verify-stride 0"
timeout 300 $SMALL > $OUT/m_small.json 2> $OUT/m_small.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/m_launches.csv $SMALL > $OUT/m_ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_seed_fast -s 3 -c 1 -o $OUT/m_prof_fast -f $SMALL > $OUT/m_ncu_full.log 2>&1
echo "ncu full rc=$?"
bash scripts/ncu_all.sh m 2000000 > $OUT/m_ncu_all.txt 2>&1; tail -20 $OUT/m_ncu_all.txt
ls -la $OUT | tail -30
