"""Ad-hoc GPU exploration: gather probe + config-1 timing.  Not part of the test suite."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import compseed_b200 as cs
from compseed_b200 import synth
from oracle import oracle_py as O

out = {}
for gran in (32, 64, 128):
    for tb in (256 << 20, 4 << 30):
        gb, gl = cs.probe_random_gather(0, tb, gran, 1 << 28, 2)
        out[f"gather_{gran}B_{tb >> 20}MiB"] = (round(gb, 1), round(gl, 2))
        print("gather", gran, tb >> 20, "MiB:", round(gb, 1), "GB/s", round(gl, 2), "Gloads/s", flush=True)

L = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
ref = synth.random_reference(L)
t = time.time(); oi = O.OracleIndex.build(ref); print("oracle index build", round(time.time() - t, 1), "s", flush=True)
bases, off, _ = synth.simulate_reads(ref, N, 150, 0.01, seed=1)
for dense in (0, 1):
    t = time.time()
    idx = cs.FMIndex.upload(oi.primary, oi.L2, oi.seq_len, oi.bwt, oi.sa, oi.sa_intv, dense_sa_intv=dense)
    print("upload dense", dense, round(time.time() - t, 2), "s", idx.device_bytes >> 20, "MiB", flush=True)
    ctx = cs.SeedContext(idx, N, int(off[-1]), 150, N * 16, N * 32, 1)
    ctx.stage(0, bases, off)
    for it in range(4):
        ctx.run_staged(0, cs.SeedOpt())
        r = ctx.wait_device(0)
        print("dense", dense, "iter", it, "kernel ms", [round(x, 3) for x in r.kernel_ms], r.counters,
              "reads/s", round(N / (sum(r.kernel_ms[:3]) * 1e-3)), flush=True)
    out[f"cfg_{L}_{N}_dense{dense}"] = dict(ms=r.kernel_ms, counters=r.counters)
    got = ctx.fetch(0)
    if dense == 0:
        t = time.time(); want = oi.seed(bases, off, n_threads=16); print("oracle seed s", round(time.time() - t, 2), want.counters)
    ok = (np.array_equal(got.mems, want.mems) and np.array_equal(got.rbeg, want.rbeg) and np.array_equal(got.mem_off, want.mem_off))
    print("parity", ok, flush=True)
    t = time.time(); r2 = cs.seed_reads(idx, bases, off, batch_reads=1 << 16, n_slots=3); dt = time.time() - t
    print("e2e seed_reads", round(dt, 3), "s", round(N / dt), "reads/s", flush=True)
    ctx.close(); idx.close()
json.dump(out, open("gpurun_out/first_run.json", "w"), indent=1)
