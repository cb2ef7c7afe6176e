#!/bin/bash
# Round-2 GPU call AI: quick ncu counters of every kernel of one step, final build, on a 200 Mbp index (a full-size index makes every
# ncu replay pass save and restore 155 GB).
set -u
OUT=gpurun_out; mkdir -p $OUT
M=smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,gpu__time_duration.sum,dram__sectors_read.sum,dram__sectors_write.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.per_cycle_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__registers_per_thread
CMD="python bench.py --ref-len 200000000 --reads 2000000 --steps 1 --warmup 1 --no-cpu --no-e2e --no-probe --verify-stride 0 --no-bsw"
timeout 200 $CMD > $OUT/ai_plain.log 2>&1 && timeout 240 ncu --metrics $M --clock-control none -k regex:'k_seed|k_walk|k_collect|k_mem_counts|k_sa_resolve|k_pack' -s 13 -c 13 --csv --log-file $OUT/ai_q.csv $CMD > $OUT/ai_ncu.log 2>&1
echo "rc=$?"
python - <<'PY'
import csv
reads=2000000
rows = [r for r in csv.reader(open("gpurun_out/ai_q.csv")) if len(r) > 10 and r[0].isdigit()]
d = {}
for r in rows: d.setdefault((int(r[0]), r[4].split('(')[0]), {})[r[12]] = float(r[14].replace(',', ''))
print(f"{'kernel':22s} {'ms':>8s} {'grid':>6s} {'regs':>5s} {'winst/read':>10s} {'lanes':>6s} {'issue%':>7s} {'warps':>6s} {'dramRd sect/read':>16s} {'dram%':>6s} {'L2hit%':>7s} {'longSB':>7s}")
for (i, k), v in sorted(d.items()):
    ms = v['gpu__time_duration.sum'] / 1e6
    print(f"{k[:22]:22s} {ms:8.3f} {v['launch__grid_size']:6.0f} {v['launch__registers_per_thread']:5.0f} {v['smsp__inst_executed.sum']/reads:10.1f} {v['smsp__thread_inst_executed_per_inst_executed.ratio']:6.2f} "
          f"{v['smsp__issue_active.avg.pct_of_peak_sustained_active']:7.2f} {v['sm__warps_active.avg.per_cycle_active']:6.2f} {v['dram__sectors_read.sum']/reads:16.1f} {v['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']:6.2f} {v['lts__t_sector_hit_rate.pct']:7.2f} {v['smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio']:7.2f}")
PY
