#!/usr/bin/env python3
"""Inter-read reuse and L2 persistence (north_star: "SST trie rebuilt as a per-batch shared-memory / L2 structure", "L2 persistence for
the hot top-of-search Occ blocks"): the cfg2 index, reads drawn from a WINDOW of the reference at high coverage (neighbouring reordered
reads overlap: the reference's SST reuses their common prefixes), position-sorted or shuffled, with or without a persisting L2 window
over the top of the k-mer table.  One JSON line: per-kernel time of a device-resident step.
usage: python scripts/r02_reuse.py [--window-mbp 20] [--reads 4000000] [--shuffle] [--persist-mb 0] [--ref-len 3100000000]"""
import argparse, json, os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
ap = argparse.ArgumentParser()
ap.add_argument("--window-mbp", type=float, default=20.0)
ap.add_argument("--reads", type=int, default=4_000_000)
ap.add_argument("--shuffle", action="store_true")
ap.add_argument("--persist-mb", type=int, default=0)
ap.add_argument("--ref-len", type=int, default=3_100_000_000)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
import torch
import compseed_b200 as cs
from compseed_b200 import synth
ref_t = synth.random_reference_torch(a.ref_len, 20261018, "cuda:0")
lo = a.ref_len // 3
hi = min(a.ref_len, lo + int(a.window_mbp * 1e6)) if a.window_mbp > 0 else a.ref_len
if a.window_mbp <= 0:
    lo = 0
bases, off, pos = synth.simulate_reads_torch(ref_t, a.reads, 150, 0.01, seed=1000, window=(lo, hi))
ref = ref_t.cpu().numpy(); del ref_t; torch.cuda.empty_cache()
if a.shuffle:
    perm = np.random.default_rng(5).permutation(a.reads)
    bases = np.ascontiguousarray(bases.reshape(a.reads, 150)[perm]).reshape(-1)
idx = cs.FMIndex.build(ref, device=0, sa_intv=1)
del ref
n = a.reads
ctx = cs.SeedContext(idx, n, int(off[-1]), 150, n * 14, n * 20, 1, cs.CtxConfig(l2_persist_mb=a.persist_mb))
ctx.stage(0, bases, off)
for _ in range(2):
    ctx.run_staged(0, cs.SeedOpt()); r = ctx.wait_device(0)
acc = None
for _ in range(a.steps):
    ctx.run_staged(0, cs.SeedOpt()); r = ctx.wait_device(0)
    km, k2 = r.kernel_ms, r.kernel_ms2
    v = np.array([km[0] + km[1] + km[2], km[6] - k2[2], km[7], k2[0], k2[1], km[1], km[2]])
    acc = v if acc is None else acc + v
acc /= a.steps
print(json.dumps({"window_mbp": a.window_mbp, "coverage": n * 150 / ((hi - lo) if a.window_mbp > 0 else a.ref_len), "reads": n, "order": "shuffled" if a.shuffle else "position-sorted",
                  "l2_persist_mb": a.persist_mb, "reads_per_s": n / (acc[0] * 1e-3), "ms": dict(zip(["step", "fast", "walk", "literal", "third", "collect", "sa"], [round(float(x), 3) for x in acc])),
                  "requests_per_read": [round(x / n, 2) for x in r.gather_requests[:5]], "mems": r.n_mems_device, "seeds": r.n_seeds_device}))
