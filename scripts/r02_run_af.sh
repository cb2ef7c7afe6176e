#!/bin/bash
# Round-2 GPU call AF (4 GPUs): the bench under torchrun at N = 4, as the driver launches it.
set -u
OUT=gpurun_out; mkdir -p $OUT
free -g | head -2 > $OUT/af_host.txt; nproc >> $OUT/af_host.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 3 --warmup 3 ) > $OUT/af_bench_n4.json 2> $OUT/af_bench_n4.err; echo "bench n4 rc=$?"; tail -4 $OUT/af_bench_n4.err
cat $OUT/af_host.txt
python - <<'PY'
import json
for l in open('gpurun_out/af_bench_n4.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N', d['n_gpus'], 'value %.1f M'%(d['value']/1e6), 'e2e %.1f M'%(d['e2e']['value']/1e6), 'chain', d['e2e']['with_chaining_on_the_gpu'].get('value'), 'host link', d['host_link']['aggregate_h2d_plus_d2h_gb_per_s'], d['host_link']['reads_per_s_this_link_allows'])
PY
