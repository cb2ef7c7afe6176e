"""(GPU box) extension kernel timing on a synthetic batch: python scripts/bsw_time.py [pairs] [ctas,...]  (COMPSEED_LIB_TAG selects a variant build)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import compseed_b200 as cs
from compseed_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
ctas = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8]
pairs, ref, qer = synth.extension_pairs_fast(n, seed=411)
ex = cs.BswExtender(0, pairs.shape[0], ref.nbytes, qer.nbytes, 256)
ex.stage(pairs, ref, qer)
for it in range(3):
    ms, cells = ex.run_staged()
print("%s rows in shared memory: %.2f ms, %.1f Gcells, %.1f GCUPS, %.1f M pairs/s" % (os.environ.get("COMPSEED_LIB_TAG", "default"), ms, cells / 1e9, cells / ms / 1e6, n / ms / 1e3), flush=True)
ex.set_rows_in_smem(False)
for c in ctas:
    ex.set_ctas_per_sm(c)
    for it in range(3):
        ms, cells = ex.run_staged()
    print("%s ctas/SM %2d: %.2f ms, %.1f Gcells, %.1f GCUPS, %.1f M pairs/s" % (os.environ.get("COMPSEED_LIB_TAG", "default"), c, ms, cells / 1e9, cells / ms / 1e6, n / ms / 1e3), flush=True)
