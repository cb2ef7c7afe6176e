#!/usr/bin/env python3
"""Full-size cross-check on the bench workload (3.1 Gbp index): the fast route (k_seed_fast + k_seed_walk +
k_seed in call mode + k_seed_r3_fast) against the literal route (k_seed in read mode + k_seed_r3), which the
parity tests pin to the oracle at small sizes.  Every mem and every seed position of every read must agree,
for several read sets and through both the device-resident and the batched host-buffer paths.
usage: python scripts/selfcheck_cfg2.py [reads_per_set] [n_sets] > profiles/r01_selfcheck_cfg2.json"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import compseed_b200 as cs
import bench

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
n_sets = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ref_len = 3_100_000_000
out = []
idx = None
for k in range(n_sets):
    args = argparse.Namespace(ref_len=ref_len, reads=n_reads, read_len=150)
    ref, bases, off = bench.make_workload(args, k, n_sets, "cuda:0")     # set k: reads from the k-th slice of the reference
    if idx is None:
        idx = cs.FMIndex.build(ref, device=0, sa_intv=1)
    del ref
    res = {}
    for mode in ("fast", "literal"):
        os.environ["CS_FAST"] = "1" if mode == "fast" else "0"
        ctx = cs.SeedContext(idx, n_reads, int(off[-1]), 150, n_reads * 14, n_reads * 20, 1)
        ctx.stage(0, bases, off)
        ctx.run_staged(0, cs.SeedOpt())
        d = ctx.wait_device(0)
        res[mode] = (ctx.fetch(0), d.counters.get("deferred_calls", 0))
        ctx.close()
    os.environ["CS_FAST"] = "1"
    batched = cs.seed_reads(idx, bases, off, cs.SeedOpt(), batch_reads=300_007, n_slots=3)   # odd batch size on purpose
    a, b = res["fast"][0], res["literal"][0]
    same = bool(np.array_equal(a.mem_off, b.mem_off) and np.array_equal(a.mems, b.mems) and np.array_equal(a.seed_off, b.seed_off) and np.array_equal(a.rbeg, b.rbeg))
    same_b = bool(np.array_equal(a.mem_off, batched.mem_off) and np.array_equal(a.mems, batched.mems) and np.array_equal(a.rbeg, batched.rbeg))
    rec = {"set": k, "reads": n_reads, "mems": int(a.mems.shape[0]), "seeds": int(a.rbeg.shape[0]), "deferred_calls_fast": int(res["fast"][1]),
           "fast_equals_literal": same, "batched_host_path_equals_device_path": same_b}
    out.append(rec)
    print(json.dumps(rec), file=sys.stderr)
print(json.dumps({"what": "scripts/selfcheck_cfg2.py: fast route vs literal route on the 3.1 Gbp index, all mems and seed positions", "results": out}, indent=1))
sys.exit(0 if all(r["fast_equals_literal"] and r["batched_host_path_equals_device_path"] for r in out) else 1)
