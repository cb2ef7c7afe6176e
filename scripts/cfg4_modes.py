#!/usr/bin/env python3
"""Repeat-rich reference (config 4): fast path on/off."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import compseed_b200 as cs
from compseed_b200 import synth
ref4 = synth.repeat_rich_reference(20_000_000, seed=41, n_segdup=2000, segdup_len=5000, n_tandem=600)
b, o, _ = synth.simulate_reads(ref4, 400_000, 150, 0.01, seed=42)
idx = cs.FMIndex.build(ref4, device=0, sa_intv=1)
n = o.shape[0] - 1
res = {}
for fast in ("1", "0"):
    os.environ["CS_FAST"] = fast
    ctx = cs.SeedContext(idx, n, int(o[-1]), 150, n * 64, n * 600, 1)
    ctx.stage(0, b, o)
    for _ in range(2):
        ctx.run_staged(0, cs.SeedOpt()); r = ctx.wait_device(0)
    got = ctx.fetch(0)
    st = ctx.debug_stats(0)
    res[fast] = got
    print(f"CS_FAST={fast}: total {r.kernel_ms[0]+r.kernel_ms[1]+r.kernel_ms[2]:.2f} ms  passes 1-2 {r.kernel_ms[4]:.2f} (fast {r.kernel_ms[6]:.2f} walk {r.kernel_ms[7]:.2f}) r3 {r.kernel_ms[5]:.2f} collect {r.kernel_ms[1]:.2f} sa {r.kernel_ms[2]:.2f}  deferred {r.counters['deferred_calls']/n:.2f}/read  ext {st[0]/n:.1f} fm {st[1]/n:.1f}")
    ctx.close()
assert np.array_equal(res["1"].mems, res["0"].mems) and np.array_equal(res["1"].rbeg, res["0"].rbeg)
print("same results")
