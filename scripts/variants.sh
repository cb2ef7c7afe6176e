#!/bin/bash
# usage: scripts/variants.sh tag1 tag2 ...   (runs a short cfg2 bench per experiment build)
mkdir -p gpurun_out
for tag in "$@"; do
  COMPSEED_LIB_TAG=$tag python bench.py --reads ${READS:-4000000} --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/var_$tag.log 2>&1
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/var_{tag}.log").read().strip().splitlines()[-1])
    sh = d["roofline"]["kernel_share_of_step"]
    print(f"{tag:10s} reads/s {d['value']/1e6:7.2f}M  ms/step {d['ms_per_step']:8.2f}  passes 1-2 {(sh['k_seed']+sh.get('k_seed_fast',0)+sh.get('k_seed_walk',0))*d['ms_per_step']:8.2f} ms  (fast {d['roofline'].get('fast_ms_per_launch', 0):7.2f}, walk {sh.get('k_seed_walk',0)*d['ms_per_step']:6.2f}, literal {sh['k_seed']*d['ms_per_step']:6.2f}, deferred {d['roofline'].get('deferred_calls_per_read', 0):5.3f} calls/read)  r3 {sh.get('k_seed_r3',0)*d['ms_per_step']:7.2f} ms  collect {sh['collect']*d['ms_per_step']:6.2f}  sa {sh['k_sa_resolve']*d['ms_per_step']:5.2f}")
except Exception as e:
    print(tag, "FAILED", e); print(open(f"gpurun_out/var_{tag}.log").read()[-800:])
PY
done
