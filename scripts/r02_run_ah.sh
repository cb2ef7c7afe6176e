#!/bin/bash
# Round-2 GPU call AH: the default bench line and the reference arm on the final build of the round.
set -u
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python bench.py ) > $OUT/ah_bench.json 2> $OUT/ah_bench.err; echo "bench rc=$?"; tail -3 $OUT/ah_bench.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > $OUT/ah_bench_ref.json 2> $OUT/ah_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/ah_bench.json').read().strip().splitlines()[-1]); r=d['roofline']
print('value %.1f M ms %.2f e2e %.1f chain %.1f parity %s frac %.3f'%(d['value']/1e6,d['ms_per_step'],d['e2e']['value']/1e6,d['e2e']['with_chaining_on_the_gpu']['value']/1e6,d['parity']['equal'],r['frac']), r['kernel_share_of_step'], d['extension_stage']['gcups'], d['extension_stage']['e2e_pairs_per_s'])
PY
