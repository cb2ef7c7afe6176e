#!/bin/bash
# Round-2 GPU call U (2 GPUs): multi-device pipeline, SAM identity on all GPUs, chains; the bench under torchrun at N = 2.
set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader > $OUT/u_gpus.txt
( time timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sam_identity.py tests/test_gpu_chain.py -m gpu -x -q ) > $OUT/u_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/u_pytest.log; tail -6 $OUT/u_pytest.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 ) > $OUT/u_bench_n2.json 2> $OUT/u_bench_n2.err; echo "bench n2 rc=$?"; tail -4 $OUT/u_bench_n2.err
python - <<'PY'
import json
for l in open('gpurun_out/u_bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N', d['n_gpus'], 'value %.1f M'%(d['value']/1e6), 'e2e %.1f M'%(d['e2e']['value']/1e6), 'host link', d['host_link'])
PY
