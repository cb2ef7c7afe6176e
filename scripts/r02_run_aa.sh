#!/bin/bash
# Round-2 GPU call AA: phase selection in k_seed_fast (CS_PHASE) -- parity tests, cfg2 4 M-read step with and without, ncu lanes.
set -u
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q ) > $OUT/aa_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/aa_pytest.log; tail -4 $OUT/aa_pytest.log
S="--reads 4000000 --steps 3 --warmup 2 --no-cpu --no-e2e --no-probe --verify-stride 0"
for tag in "" nophase; do
  COMPSEED_LIB_TAG=$tag timeout 300 python bench.py $S > $OUT/aa_small_$tag.json 2> $OUT/aa_small_$tag.err; echo "small '$tag' rc=$?"
done
python - <<'PY'
import json
for tag in ("","nophase"):
    d=json.loads(open(f'gpurun_out/aa_small_{tag}.json').read().strip().splitlines()[-1]); r=d['roofline']
    print(repr(tag),'value %.1f M'%(d['value']/1e6),'ms %.2f'%d['ms_per_step'],{k:round(v['ms_per_step'],2) for k,v in r['all_kernels'].items()})
PY
