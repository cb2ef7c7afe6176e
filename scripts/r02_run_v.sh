#!/bin/bash
# Round-2 GPU call V (final evidence of the round): the whole GPU test suite, smoke(), the default bench line as the driver runs it, the
# reference arm, the launch list of two reduced steps, full ncu captures of k_seed_fast and k_seed_walk, cfg4.
set -u
OUT=gpurun_out; mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $OUT/v_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/v_pytest.log; tail -6 $OUT/v_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/v_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/v_smoke.log
( time timeout 900 python bench.py ) > $OUT/v_bench.json 2> $OUT/v_bench.err; echo "bench rc=$?"; tail -3 $OUT/v_bench.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > $OUT/v_bench_ref.json 2> $OUT/v_bench_ref.err; echo "ref rc=$?"
timeout 300 python scripts/r02_cfg4.py > $OUT/v_cfg4.json 2> $OUT/v_cfg4.err; echo "cfg4 rc=$?"
SMALL="python bench.py --reads 2000000 --steps 2 --warmup 3 --no-cpu --no-e2e --no-probe --verify-stride 0"
timeout 300 $SMALL > $OUT/v_small.json 2> $OUT/v_small.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_seed|k_walk|k_collect|k_mem_counts|k_sa_resolve|k_pack|DeviceScan|k_compact' --csv --log-file $OUT/v_launches.csv $SMALL > $OUT/v_ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_seed_fast|k_seed_walk' -s 6 -c 2 -o $OUT/v_prof_seed -f $SMALL > $OUT/v_ncu_full.log 2>&1
echo "ncu full rc=$?"
timeout 120 python scripts/bsw_time.py 2000000 8 > $OUT/v_bsw_time.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bsw_extend_smem -s 2 -c 1 -o $OUT/v_prof_bsw -f python scripts/bsw_time.py 2000000 8 > $OUT/v_ncu_bsw.log 2>&1
echo "ncu bsw rc=$?"; cat $OUT/v_bsw_time.log
ls -la $OUT | grep " v_"
