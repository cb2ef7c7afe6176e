#!/bin/bash
# Round-2 GPU call W (2 GPUs): the bench under torchrun at N = 2, as the driver launches it.
set -u
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 ) > $OUT/w_bench_n2.json 2> $OUT/w_bench_n2.err; echo "bench n2 rc=$?"; tail -4 $OUT/w_bench_n2.err
python - <<'PY'
import json
for l in open('gpurun_out/w_bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N', d['n_gpus'], 'value %.1f M'%(d['value']/1e6), 'e2e %.1f M'%(d['e2e']['value']/1e6), 'chain', d['e2e']['with_chaining_on_the_gpu'], 'host link', d['host_link'])
PY
