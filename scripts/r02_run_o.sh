#!/bin/bash
# Round-2 GPU call O: default bench with the extension-stage leg; inter-read reuse / L2 persistence experiment (timings + L2/L1 hit rates
# of k_seed_fast by ncu); launch list of one reduced bench step; full ncu capture of k_seed_fast (repeat-length build) and of the BSW kernel.
set -u
OUT=gpurun_out; mkdir -p $OUT
( time timeout 600 python bench.py ) > $OUT/o_bench.json 2> $OUT/o_bench.err; echo "bench rc=$?"; tail -3 $OUT/o_bench.err
for v in "" "--shuffle" "--persist-mb 32" "--window-mbp 0" "--window-mbp 0 --persist-mb 32"; do
  timeout 200 python scripts/r02_reuse.py $v >> $OUT/o_reuse.jsonl 2>> $OUT/o_reuse.err; echo "reuse '$v' rc=$?"
done
cat $OUT/o_reuse.jsonl
M=gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,dram__sectors_read.sum,lts__t_requests_srcunit_tex_op_read.sum
for v in "" "--shuffle" "--persist-mb 32"; do
  tag=$(echo "sorted$v" | tr -d ' -')
  timeout 300 ncu --metrics $M --clock-control none -k regex:k_seed_fast -s 2 -c 1 --csv --log-file $OUT/o_reuse_ncu_$tag.csv python scripts/r02_reuse.py $v --steps 1 > $OUT/o_reuse_ncu_$tag.log 2>&1; echo "ncu reuse $tag rc=$?"
done
SMALL="python bench.py --reads 2000000 --steps 2 --warmup 3 --no-cpu --no-e2e --no-probe --verify-stride 0"
timeout 300 $SMALL > $OUT/o_small.json 2> $OUT/o_small.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_seed|k_collect|k_mem_counts|k_sa_resolve|k_pack|DeviceScan|k_compact' --csv --log-file $OUT/o_launches.csv $SMALL > $OUT/o_ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_seed_fast -s 3 -c 1 -o $OUT/o_prof_fast -f $SMALL > $OUT/o_ncu_full.log 2>&1
echo "ncu full rc=$?"
timeout 120 python scripts/bsw_time.py 2000000 8 > $OUT/o_bsw_time.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bsw_extend -s 2 -c 1 -o $OUT/o_prof_bsw -f python scripts/bsw_time.py 2000000 8 > $OUT/o_ncu_bsw.log 2>&1
echo "ncu bsw rc=$?"; cat $OUT/o_bsw_time.log
ls -la $OUT | tail -30
