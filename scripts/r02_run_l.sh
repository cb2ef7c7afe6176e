#!/bin/bash
# Round-2 GPU call L: extension kernel v2 (16-bit cells, next-column prefetch, shape-sorted pairs): parity + timing by CTAs per SM.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_bsw.py -m gpu -q -x > $OUT/l_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/l_pytest.log; tail -5 $OUT/l_pytest.log
timeout 600 python - > $OUT/l_bsw_time.log 2>&1 <<'PY'
import time, numpy as np, os
import compseed_b200 as cs
from compseed_b200 import synth
pairs, ref, qer = synth.extension_pairs_fast(2_000_000, seed=411)
ex = cs.BswExtender(0, pairs.shape[0], ref.nbytes, qer.nbytes, 256)
ex.stage(pairs, ref, qer)
for c in (2, 3, 4, 6, 8, 12, 16):
    ex.set_ctas_per_sm(c)
    for it in range(3):
        ms, cells = ex.run_staged()
    print("ctas/SM %2d: %.2f ms, %.1f Gcells, %.1f GCUPS, %.1f M pairs/s" % (c, ms, cells / 1e9, cells / ms / 1e6, pairs.shape[0] / ms / 1e3))
PY
cat $OUT/l_bsw_time.log
