#!/bin/bash
# Round-2 GPU call K: extension kernel (cs_bsw_*) parity tests and a first timing.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_bsw.py -m gpu -q -x > $OUT/k_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/k_pytest.log; tail -15 $OUT/k_pytest.log
timeout 600 python - > $OUT/k_bsw_time.log 2>&1 <<'PY'
import time, numpy as np, os
import compseed_b200 as cs
from compseed_b200 import synth
pairs, ref, qer = synth.extension_pairs_fast(2_000_000, seed=411)
ex = cs.BswExtender(0, pairs.shape[0], ref.nbytes, qer.nbytes, 256)
ex.stage(pairs, ref, qer)
for it in range(4):
    ms, cells = ex.run_staged()
    print("staged: %.2f ms, %.1f Gcells, %.1f GCUPS, %.1f M pairs/s" % (ms, cells / 1e9, cells / ms / 1e6, pairs.shape[0] / ms / 1e3))
t = time.time(); ex.extend(pairs, ref, qer); dt = time.time() - t
print("host buffers: %.1f ms, %.1f M pairs/s" % (dt * 1e3, pairs.shape[0] / dt / 1e6))
PY
cat $OUT/k_bsw_time.log
