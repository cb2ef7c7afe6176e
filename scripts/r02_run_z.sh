#!/bin/bash
# Round-2 GPU call Z: extension-stage host path on the host's threads (parity tests), and the default bench line of the final build.
set -u
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_gpu_bsw.py -m gpu -x -q ) > $OUT/z_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/z_pytest.log; tail -4 $OUT/z_pytest.log
( time timeout 900 python bench.py ) > $OUT/z_bench.json 2> $OUT/z_bench.err; echo "bench rc=$?"; tail -3 $OUT/z_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/z_bench.json').read().strip().splitlines()[-1])
print('value %.1f M ms %.2f e2e %.1f parity %s'%(d['value']/1e6,d['ms_per_step'],d['e2e']['value']/1e6,d['parity']['equal']), d['extension_stage']['e2e_pairs_per_s'], d['extension_stage']['parity']['equal'])
PY
