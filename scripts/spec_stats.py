#!/usr/bin/env python3
"""Event counters of k_seed on the bench workload (needs a -DCS_STATS build:
CS_DEFS=-DCS_STATS CS_TAG=stats python -m compseed_b200.build).
usage: COMPSEED_LIB_TAG=stats python scripts/spec_stats.py [n_reads] [ref_len]"""
import argparse, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import compseed_b200 as cs
import bench

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ref_len = int(sys.argv[2]) if len(sys.argv) > 2 else 3_100_000_000
NAMES = ["warp trips", "lane kt lookups", "lane FM extends", "lane text-path trips", "calls", "spec calls", "spec success",
         "spec stop (nothing pushable)", "spec abort: stop", "spec abort: K-mer present", "spec abort: probe not applicable",
         "lane LF trips", "lane bookkeeping entries", "filter requests", "jump rejects", "lane text-compare trips"]
args = argparse.Namespace(ref_len=ref_len, reads=n_reads, read_len=150)
ref, bases, off = bench.make_workload(args, 0, 1, "cuda:0")
idx = cs.FMIndex.build(ref, device=0, sa_intv=1)
ctx = cs.SeedContext(idx, n_reads, int(off[-1]), 150, n_reads * 14, n_reads * 20, 1)
ctx.stage(0, bases, off)
for _ in range(2):
    ctx.run_staged(0, cs.SeedOpt())
    r = ctx.wait_device(0)
st = ctx.debug_stats(0)
print(f"reads {n_reads}  passes 1-2 {r.kernel_ms[4]:.2f} ms (pack + fast {r.kernel_ms[6]:.2f} ms, walk {r.kernel_ms[7]:.2f} ms)  r3 {r.kernel_ms[5]:.2f} ms  collect {r.kernel_ms[1]:.2f} ms")
print(f"  per read: ext queries {st[0]/n_reads:.1f}  FM extends (k_seed + r3) {st[1]/n_reads:.1f}  filter probes {st[3]/n_reads:.1f}")
for k, name in enumerate(NAMES):
    print(f"  {name:36s} {int(st[4 + k]):14d}  {st[4 + k] / n_reads:9.3f} / read")
print(f"  deferred to the literal kernel: {int(st[20])} reads ({100.0 * st[20] / n_reads:.2f} %), grids fast/literal {int(st[21])}/{int(st[22])}")
FAST = ["warp iterations", "calls", "pass-2 calls", "done: nothing pushable", "defer: pass-2 call needs a list", "defer: L not pushed",
        "L followed backward through the FM-index first", "defer: > 64 backward FM steps", "defer: K-mer present", "defer: probe not applicable",
        "L resolved", "mems emitted"]
print("k_seed_fast:")
for k, name in enumerate(FAST):
    print(f"  {name:48s} {int(st[24 + k]):14d}  {st[24 + k] / n_reads:9.3f} / read")
print(json.dumps({"reads": n_reads, "stats": [int(v) for v in st]}))
