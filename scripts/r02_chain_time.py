#!/usr/bin/env python3
"""Chaining stage (mem_chain + mem_chain_flt on the device) on the cfg2 index: a device-resident step with chaining on;
run under `ncu --metrics gpu__time_duration.sum -k regex:k_chain` for the kernel times.  usage: python scripts/r02_chain_time.py [reads]"""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import numpy as np
import compseed_b200 as cs
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
args = argparse.Namespace(ref_len=3_100_000_000, reads=n, read_len=150)
ref, bases, off = bench.make_workload(args, 0, 1, "cuda:0")
idx = cs.FMIndex.build(ref, device=0, sa_intv=1)
del ref
ctx = cs.SeedContext(idx, n, int(off[-1]), 150, n * 14, n * 20, 1)
ctx.set_chaining(bench.contig_lens(args.ref_len))
ctx.stage(0, bases, off)
for it in range(3):
    t0 = time.perf_counter()
    ctx.run_staged(0, cs.SeedOpt()); r = ctx.wait_device(0)
    dt = time.perf_counter() - t0
    km = r.kernel_ms
    print("step %d: wall %.2f ms, seeding+collect+sa kernels %.2f ms => chaining + the rest %.2f ms" % (it, dt * 1e3, km[0] + km[1] + km[2], dt * 1e3 - (km[0] + km[1] + km[2])), flush=True)
