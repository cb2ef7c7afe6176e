#!/bin/bash
# Round-2 GPU call AJ: resident CTAs of the literal kernel in call mode (cs_ctx_config_t.lit_ctas_per_sm 1 vs default) on cfg4 and cfg2.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python - > $OUT/aj_cfg4.txt 2>&1 <<'PY'
import json, os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import compseed_b200 as cs
from compseed_b200 import synth
ref4 = synth.repeat_rich_reference(20_000_000, seed=41, n_segdup=2000, segdup_len=5000, n_tandem=600)
b, o, _ = synth.simulate_reads(ref4, 400_000, 150, 0.01, seed=42)
idx = cs.FMIndex.build(ref4, device=0, sa_intv=1)
n = o.shape[0] - 1
for name, cfg in (("default", cs.CtxConfig()), ("lit1", cs.CtxConfig(lit_ctas_per_sm=1))):
    ctx = cs.SeedContext(idx, n, int(o[-1]), 150, n * 64, n * 600, 1, cfg)
    ctx.stage(0, b, o)
    for _ in range(3):
        ctx.run_staged(0, cs.SeedOpt()); r = ctx.wait_device(0)
    km, k2 = r.kernel_ms, r.kernel_ms2
    tot = km[0] + km[1] + km[2]
    print(name, "cfg4 %.1f M reads/s  total %.2f ms  literal %.2f ms" % (n / tot / 1e3, tot, k2[0]), flush=True)
    ctx.close()
PY
cat $OUT/aj_cfg4.txt
S="--reads 4000000 --steps 3 --warmup 2 --no-cpu --no-e2e --no-probe --verify-stride 0 --no-bsw"
timeout 200 python bench.py $S --lit-ctas 1 > $OUT/aj_small_lit1.json 2> $OUT/aj_small_lit1.err; echo "small rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/aj_small_lit1.json').read().strip().splitlines()[-1]); r=d['roofline']
print('cfg2 lit1: value %.1f M'%(d['value']/1e6),{k:round(v['ms_per_step'],2) for k,v in r['all_kernels'].items()})
PY
