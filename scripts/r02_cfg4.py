#!/usr/bin/env python3
"""BASELINE.json configs[3] (repeat-rich 20 Mbp reference, 400 k x 150 bp reads) and a 4 M-read slice of configs[1]:
device-resident step time per kernel for the library selected by COMPSEED_LIB_TAG.  usage: python scripts/r02_cfg4.py [cfg2]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import compseed_b200 as cs
from compseed_b200 import synth
ref4 = synth.repeat_rich_reference(20_000_000, seed=41, n_segdup=2000, segdup_len=5000, n_tandem=600)
b, o, _ = synth.simulate_reads(ref4, 400_000, 150, 0.01, seed=42)
idx = cs.FMIndex.build(ref4, device=0, sa_intv=1)
n = o.shape[0] - 1
out = {"tag": os.environ.get("COMPSEED_LIB_TAG", "")}
for name, cfg in (("default", cs.CtxConfig()), ("lit2", cs.CtxConfig(lit_ctas_per_sm=2)), ("nofast", cs.CtxConfig(use_fast=0))):
    ctx = cs.SeedContext(idx, n, int(o[-1]), 150, n * 64, n * 600, 1, cfg)
    ctx.stage(0, b, o)
    for _ in range(3):
        ctx.run_staged(0, cs.SeedOpt()); r = ctx.wait_device(0)
    km, k2 = r.kernel_ms, r.kernel_ms2
    tot = km[0] + km[1] + km[2]
    out[name] = {"reads_per_s": n / (tot * 1e-3), "ms": tot, "fast": km[6] - k2[2], "walk": km[7], "literal": k2[0], "third": k2[1], "collect": km[1], "sa": km[2],
                 "deferred_per_read": r.counters["deferred_calls"] / n, "mems_per_read": r.n_mems_device / n, "seeds_per_read": r.n_seeds_device / n}
    ctx.close()
print(json.dumps(out))
