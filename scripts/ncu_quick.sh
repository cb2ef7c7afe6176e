#!/bin/bash
# quick ncu counters of k_seed / k_seed_r3 for the current build; env passes through (e.g. CS_KMER_TABLE_DEPTH)
M=smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,gpu__time_duration.sum,dram__sectors_read.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.per_cycle_active,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
TAG=$1
CMD="python bench.py --reads 2000000 --steps 1 --warmup 1 --no-cpu --no-e2e --no-probe --verify-stride 0"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && ncu --metrics $M --clock-control none -k regex:k_seed -s 2 -c 2 --csv --log-file gpurun_out/q_$TAG.csv $CMD > gpurun_out/ncu_$TAG.log 2>&1
python - "$TAG" <<'PY'
import csv, sys
tag = sys.argv[1]
rows = [r for r in csv.reader(open(f"gpurun_out/q_{tag}.csv")) if len(r) > 10 and r[0].isdigit()]
d = {}
for r in rows: d.setdefault(r[4].split('(')[0], {})[r[12]] = r[14]
for k, v in d.items():
    print(tag, k)
    for m, x in sorted(v.items()): print(f"   {m:85s} {x}")
PY
