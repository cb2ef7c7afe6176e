#!/bin/bash
# short cfg2 bench for several top-of-search table depths
mkdir -p gpurun_out
for d in "$@"; do
  CS_KMER_TABLE_DEPTH=$d python bench.py --reads ${READS:-4000000} --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/kt_$d.log 2>&1
  python - "$d" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/kt_{tag}.log").read().strip().splitlines()[-1])
    sh = d["roofline"]["kernel_share_of_step"]; c = d["counters"]
    print(f"depth {tag:3s} reads/s {d['value']/1e6:7.2f}M  ms/step {d['ms_per_step']:8.2f}  k_seed {sh['k_seed']*d['ms_per_step']:8.2f} ms  r3 {sh.get('k_seed_r3',0)*d['ms_per_step']:7.2f} ms  collect {sh['collect']*d['ms_per_step']:6.2f}  calls/queries {c['ext_calls']/c['ext_queries']:.3f}  idx {d['config']['index_bytes_per_gpu']/1e9:.1f} GB")
except Exception as e:
    print(tag, "FAILED", e); print(open(f"gpurun_out/kt_{tag}.log").read()[-1500:])
PY
done
