#!/bin/bash
# Round-2 GPU call Y: k_collect_rows rewrite -- parity tests (incl. the repeat-rich sets with -c 50/500), timing.
set -u
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_chain.py -m gpu -x -q ) > $OUT/y_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/y_pytest.log; tail -5 $OUT/y_pytest.log
S="--reads 4000000 --steps 3 --warmup 2 --no-cpu --no-e2e --no-probe --verify-stride 0"
timeout 300 python bench.py $S > $OUT/y_small.json 2> $OUT/y_small.err; echo "small rc=$?"
timeout 300 python scripts/r02_cfg4.py > $OUT/y_cfg4.json 2> $OUT/y_cfg4.err; echo "cfg4 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/y_small.json').read().strip().splitlines()[-1]); r=d['roofline']
print('value %.1f M'%(d['value']/1e6),'ms %.2f'%d['ms_per_step'],{k:round(v['ms_per_step'],2) for k,v in r['all_kernels'].items()})
c=json.load(open('gpurun_out/y_cfg4.json'))['default']; print('cfg4', {k:(round(v,2) if isinstance(v,float) else v) for k,v in c.items()})
PY
