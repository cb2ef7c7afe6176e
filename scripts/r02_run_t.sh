#!/bin/bash
# Round-2 GPU call T: walk tasks in descending order of expected length -- parity tests, cfg2 4 M-read step, cfg4.
set -u
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q ) > $OUT/t_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/t_pytest.log; tail -6 $OUT/t_pytest.log
S="--reads 4000000 --steps 3 --warmup 2 --no-cpu --no-e2e --no-probe --verify-stride 0"
timeout 300 python bench.py $S > $OUT/t_small.json 2> $OUT/t_small.err; echo "small rc=$?"
timeout 300 python scripts/r02_cfg4.py > $OUT/t_cfg4.json 2> $OUT/t_cfg4.err; echo "cfg4 rc=$?"; cat $OUT/t_cfg4.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t_small.json').read().strip().splitlines()[-1]); r=d['roofline']
print('value %.1f M'%(d['value']/1e6),'ms %.2f'%d['ms_per_step'],{k:round(v['ms_per_step'],2) for k,v in r['all_kernels'].items()})
PY
