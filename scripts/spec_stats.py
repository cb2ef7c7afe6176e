#!/usr/bin/env python3
"""Event counters of the seeding kernels on the bench workload (needs a -DCS_STATS build:
CS_DEFS=-DCS_STATS CS_TAG=stats python -m compseed_b200.build).
usage: COMPSEED_LIB_TAG=stats python scripts/spec_stats.py [n_reads] [ref_len] [repeat]"""
import argparse, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import compseed_b200 as cs
from compseed_b200 import synth
import bench

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ref_len = int(sys.argv[2]) if len(sys.argv) > 2 else 3_100_000_000
repeat = len(sys.argv) > 3 and sys.argv[3] == "repeat"
NAMES = ["warp trips", "lane kt lookups", "lane FM extends", "lane text-path trips", "calls", "literal tasks from k_seed_fast: list call with > 6 entries / d >= 256", "... L not unique with d >= 256 / unique deeper than K with the K-mer present",
         "... scratch full / > 6 shallow entries", "literal tasks punted by k_seed_walk", "second-pass follow-ups of what k_seed_walk found", "pass-2 answered at the SMEM with > 6 entries",
         "lane LF trips", "lane bookkeeping entries", "filter requests", "jump rejects", "lane text-compare trips"]
if repeat:
    ref = synth.repeat_rich_reference(ref_len, seed=41, n_segdup=2000, segdup_len=5000, n_tandem=600)
    bases, off, _ = synth.simulate_reads(ref, n_reads, 150, 0.01, seed=42)
else:
    args = argparse.Namespace(ref_len=ref_len, reads=n_reads, read_len=150)
    ref, bases, off = bench.make_workload(args, 0, 1, "cuda:0")
idx = cs.FMIndex.build(ref, device=0, sa_intv=1)
ctx = cs.SeedContext(idx, n_reads, int(off[-1]), 150, n_reads * 64, n_reads * 600 if repeat else n_reads * 20, 1)
ctx.stage(0, bases, off)
for _ in range(2):
    ctx.run_staged(0, cs.SeedOpt())
    r = ctx.wait_device(0)
st = ctx.debug_stats(0)
km, k2 = r.kernel_ms, r.kernel_ms2
print(f"reads {n_reads}  pack {k2[2]:.2f}  fast {km[6]-k2[2]:.2f}  walk {km[7]:.2f}  literal {k2[0]:.2f}  third pass (overlapped) {k2[1]:.2f}  collect {km[1]:.2f}  sa {km[2]:.2f} ms")
print(f"  per read: ext queries {st[0]/n_reads:.1f}  FM extends {st[1]/n_reads:.1f}  two-sector {st[2]/n_reads:.1f}  filter probes {st[3]/n_reads:.1f}  requests fast/walk/lit/r3 {[round(x/n_reads,1) for x in r.gather_requests[:4]]}")
print(f"  deferred calls {int(st[20])} ({st[20]/n_reads:.3f}/read), of them run by the literal kernel {int(st[23])} ({st[23]/n_reads:.4f}/read); grids fast/literal {int(st[21])}/{int(st[22])}")
print("k_seed (literal, call mode):")
for k, name in enumerate(NAMES):
    print(f"  {name:36s} {int(st[4 + k]):14d}  {st[4 + k] / n_reads:9.3f} / read   {st[4 + k] / max(1, st[23]):10.2f} / literal call")
FAST = ["warp iterations", "calls", "pass-2 calls", "done: nothing pushable", "defer: pass-2 call with a list", "defer: L not pushed",
        "defer: L not unique", "diagonal speculation: taken", "K-mer present at the failing position", "diagonal speculation: failed", "L resolved", "mems emitted",
        "pass 2 answered at the SMEM: nothing", "pass 2 answered at the SMEM: walk task", "pass 2 answered at the SMEM: literal task", "pass 2 at the SMEM: not decidable"]
print("k_seed_fast:")
for k, name in enumerate(FAST):
    print(f"  {name:48s} {int(st[24 + k]):14d}  {st[24 + k] / n_reads:9.3f} / read")
print(json.dumps({"reads": n_reads, "stats": [int(v) for v in st], "kernel_ms": km, "kernel_ms2": k2}))
