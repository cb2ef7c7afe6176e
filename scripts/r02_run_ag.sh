#!/bin/bash
# Round-2 GPU call AG: third-pass kernel with the repeat-length byte -- parity tests, cfg2 4 M-read step.
set -u
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q ) > $OUT/ag_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/ag_pytest.log; tail -4 $OUT/ag_pytest.log
S="--reads 4000000 --steps 3 --warmup 2 --no-cpu --no-e2e --no-probe --verify-stride 0"
timeout 300 python bench.py $S > $OUT/ag_small.json 2> $OUT/ag_small.err; echo "small rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/ag_small.json').read().strip().splitlines()[-1]); r=d['roofline']
print('value %.1f M'%(d['value']/1e6),'ms %.2f'%d['ms_per_step'],{k:round(v['ms_per_step'],2) for k,v in r['all_kernels'].items()}, {k:v['requests_per_read'] for k,v in r['all_kernels'].items() if v['requests_per_read']})
PY
