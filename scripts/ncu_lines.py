#!/usr/bin/env python3
"""Per-source-line instruction counts and lane efficiency from an ncu report captured with --import-source on.
usage: ncu_lines.py report.ncu-rep kernel_name [top_n]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern, "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = []
fname = None; hdr = None
for rec in csv.reader(io.StringIO(out)):
    if not rec: continue
    if rec[0] == "File Path": fname = rec[1].split("/")[-1]; continue
    if rec[0] == "Function Name": continue
    if rec[0] == "Line No": hdr = rec; continue
    if hdr is None or rec[0] == "": continue
    d = dict(zip(hdr, rec))
    try:
        rows.append((fname, int(rec[0]), rec[1].strip()[:90], int(d["Instructions Executed"]), int(d["Thread Instructions Executed"]), int(d["# Samples"])))
    except (ValueError, KeyError):
        pass
ti = sum(r[3] for r in rows); tt = sum(r[4] for r in rows); ts = sum(r[5] for r in rows)
print(f"total warp-inst {ti:,}  thread-inst {tt:,}  avg lanes {tt/max(ti,1):.2f}  samples {ts:,}")
rows.sort(key=lambda r: -r[3])
print(f"{'file:line':28s} {'winst%':>7s} {'lanes':>6s} {'smpl%':>6s}  source")
for f, ln, src, wi, th, sm in rows[:top]:
    print(f"{f+':'+str(ln):28s} {100*wi/ti:7.2f} {th/max(wi,1):6.2f} {100*sm/max(ts,1):6.2f}  {src}")
