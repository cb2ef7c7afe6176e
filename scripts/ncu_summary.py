#!/usr/bin/env python3
"""Summary of one kernel of an ncu report (ncu --set full ... -o X): the metrics the roofline discussion uses, the hottest source
lines, and (optionally) the traffic json bench.py reads for roofline.traffic.
usage: ncu_summary.py report.ncu-rep kernel_name reads_in_launch out.txt [traffic.json] [header text]"""
import csv, io, json, subprocess, sys
rep, kern, reads, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
tj = sys.argv[5] if len(sys.argv) > 5 and sys.argv[5] else None
header = sys.argv[6] if len(sys.argv) > 6 else ""
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__sectors_read.sum", "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_requests_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_ltcfabric.sum",
        "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
lines = [f"# {header}" if header else f"# {kern}", f"# ncu -i {rep.split('/')[-1]} --page raw --csv --kernel-name {kern}; {reads} reads in the launch", ""]
for k in WANT:
    if k in m:
        lines.append(f"{k:110s} {m[k][0]:>18s} {m[k][1]}")
def num(k):
    return float(m[k][0].replace(",", "")) if k in m else 0.0
def to_bytes(k):
    v, u = num(k), m[k][1].lower() if k in m else ""
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)
dram = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
l2req = num("lts__t_requests_srcunit_tex_op_read.sum")
lines += ["", f"per read: DRAM bytes (read + write) {dram / reads:.1f}, DRAM sectors read {num('dram__sectors_read.sum') / reads:.1f}, L2 read requests {l2req / reads:.1f}, "
              f"L2 read sectors {num('lts__t_sectors_srcunit_tex_op_read.sum') / reads:.1f}, warp instructions {num('smsp__inst_executed.sum') / reads:.1f}"]
src = subprocess.run(["python3", __file__.replace("ncu_summary.py", "ncu_lines.py"), rep, kern, "25"], capture_output=True, text=True).stdout
lines += ["", "hottest source lines (warp instructions):", src]
open(out, "w").write("\n".join(lines) + "\n")
if tj:
    json.dump({"kernel": kern, "capture": out.split("/")[-1], "reads_in_capture": reads, "dram_bytes_per_read": dram / reads,
               "l2_read_requests_per_read": l2req / reads, "dram_sectors_read_per_read": num("dram__sectors_read.sum") / reads}, open(tj, "w"), indent=1)
print("\n".join(lines[:45]))
