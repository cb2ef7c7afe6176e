#!/bin/bash
# Round-2 GPU call AE: more entries / longer walks per task in k_seed_walk (CS_WALK_MAX 18, CS_WALK_STEPS 250) on cfg2 and cfg4.
set -u
OUT=gpurun_out; mkdir -p $OUT
S="--reads 4000000 --steps 3 --warmup 2 --no-cpu --no-e2e --no-probe --verify-stride 0"
for tag in "" wm18 wm18s; do
  COMPSEED_LIB_TAG=$tag timeout 300 python bench.py $S > $OUT/ae_small_$tag.json 2> $OUT/ae_small_$tag.err; echo "small '$tag' rc=$?"
  COMPSEED_LIB_TAG=$tag timeout 300 python scripts/r02_cfg4.py > $OUT/ae_cfg4_$tag.json 2> $OUT/ae_cfg4_$tag.err; echo "cfg4 '$tag' rc=$?"
done
python - <<'PY'
import json
for tag in ("","wm18","wm18s"):
    d=json.loads(open(f'gpurun_out/ae_small_{tag}.json').read().strip().splitlines()[-1]); r=d['roofline']
    print(repr(tag),'cfg2 %.1f M'%(d['value']/1e6),{k:round(v['ms_per_step'],2) for k,v in r['all_kernels'].items()})
    c=json.load(open(f'gpurun_out/ae_cfg4_{tag}.json'))['default']; print('     cfg4 %.1f M'%(c['reads_per_s']/1e6), {k:round(c[k],2) for k in ('fast','walk','literal','third')})
PY
