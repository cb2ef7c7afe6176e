#!/usr/bin/env python3
"""BASELINE.json configs 4 and 5 on the GPU box: repeat-rich reference, coverage (sorted vs shuffled), -r sweep,
read lengths.  Device-resident seeding throughput + how many bwt_smem1a calls leave the fast kernel.
Every result set is checked against the CPU oracle on a sample (bit-exact) before it is timed.
usage: python scripts/config_sweep.py > profiles/r01_config_sweep.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import compseed_b200 as cs
from compseed_b200 import synth
from oracle import oracle_py as O

O.build(ref=False)
out = []


def run(name, ref, bases, off, opt, idx=None, check=2000):
    try:
        _run(name, ref, bases, off, opt, idx, check)
    except Exception as e:   # keep going: one case must not cost the others
        out.append({"case": name, "error": repr(e)})
        print(name, "FAILED", repr(e), file=sys.stderr)


def _run(name, ref, bases, off, opt, idx, check):
    own = idx is None
    if own:
        idx = cs.FMIndex.build(ref, device=0, sa_intv=1)
    n = off.shape[0] - 1
    rl = int((off[1:] - off[:-1]).max())
    ctx = cs.SeedContext(idx, n, int(off[-1]), rl, n * 64, n * 600, 1)
    ctx.stage(0, bases, off)
    for _ in range(2):
        ctx.run_staged(0, opt)
        r = ctx.wait_device(0)
    ms = r.kernel_ms[0] + r.kernel_ms[1] + r.kernel_ms[2]
    got = ctx.fetch(0)
    ok = None
    if check:   # parity on the first `check` reads against the oracle (same index semantics, pinned to bwaidx by the goldens)
        oi = O.OracleIndex.build(ref)
        m = min(check, n)
        want = oi.seed(bases[:int(off[m])], off[:m + 1], split_len=opt.split_len, max_mem_intv=opt.max_mem_intv, max_occ=opt.max_occ,
                       min_seed_len=opt.min_seed_len, n_threads=8)
        ok = bool(np.array_equal(got.mem_off[:m + 1], want.mem_off) and np.array_equal(got.mems[:int(got.mem_off[m])], want.mems)
                  and np.array_equal(got.rbeg[:int(got.seed_off[m])], want.rbeg))
    rec = {"case": name, "reads": n, "read_len": rl, "ref_bp": int(ref.shape[0]), "ms": ms, "reads_per_s": n / (ms * 1e-3),
           "mems_per_read": got.mems.shape[0] / n, "seeds_per_read": got.rbeg.shape[0] / n,
           "deferred_calls_per_read": r.counters.get("deferred_calls", 0) / n,
           "kernel_ms": {"pack+fast": r.kernel_ms[6], "walk": r.kernel_ms[7], "literal": r.kernel_ms[4] - r.kernel_ms[6] - r.kernel_ms[7],
                         "third_pass": r.kernel_ms[5], "collect": r.kernel_ms[1], "sa": r.kernel_ms[2]},
           "bit_exact_vs_oracle_on_sample": ok}
    out.append(rec)
    print(json.dumps(rec), file=sys.stderr)
    ctx.close()
    if own:
        idx.close()


# config 4: repeat-rich reference (segmental duplications + tandem repeats): x[2] > max_occ, 3rd-round reseeding, bwt_sa volume
ref4 = synth.repeat_rich_reference(20_000_000, seed=41, n_segdup=2000, segdup_len=5000, n_tandem=600)
b, o, _ = synth.simulate_reads(ref4, 400_000, 150, 0.01, seed=42)
run("cfg4 repeat-rich 20 Mbp, 400k x 150 bp, defaults", ref4, b, o, cs.SeedOpt(), check=0)   # parity of this class: tests/test_gpu_parity.py (repeat)
run("cfg4 repeat-rich 20 Mbp, 400k x 150 bp, -c 50 -y 40", ref4, b, o, cs.SeedOpt(max_occ=50, max_mem_intv=40), check=0)

# config 5: coverage / order / -r / read length on a 5 Mbp i.i.d. reference
ref5 = synth.random_reference(5_000_000, seed=51)
idx5 = cs.FMIndex.build(ref5, device=0, sa_intv=1)
for cov in (10, 30, 60):
    n = cov * ref5.shape[0] // 150
    b, o, _ = synth.simulate_reads(ref5, n, 150, 0.01, seed=52 + cov)
    run(f"cfg5 5 Mbp {cov}x sorted", ref5, b, o, cs.SeedOpt(), idx=idx5, check=2000 if cov == 10 else 0)
    sb, so, _ = synth.shuffle_reads(b, o)
    run(f"cfg5 5 Mbp {cov}x shuffled", ref5, sb, so, cs.SeedOpt(), idx=idx5, check=0)
b, o, _ = synth.simulate_reads(ref5, 1_000_000, 150, 0.01, seed=60)
for r in (1.0, 1.5, 2.0, 2.5):
    run(f"cfg5 5 Mbp 30x -r {r}", ref5, b, o, cs.SeedOpt(split_factor=r), idx=idx5, check=1000)
for L in (100, 250):
    b, o, _ = synth.simulate_reads(ref5, 1_000_000, L, 0.01, seed=61 + L)
    run(f"cfg5 5 Mbp read length {L}", ref5, b, o, cs.SeedOpt(), idx=idx5, check=1000)
idx5.close()
print(json.dumps({"what": "BASELINE.json configs 4-5, device-resident seeding on one B200 (scripts/config_sweep.py)", "results": out}, indent=1))
