"""Random-sector gather sweep: table size, loads in flight per thread, L2 fetch granularity."""
import json
import sys
sys.path.insert(0, ".")
import compseed_b200 as cs
out = []
for fg in (0, 32):
    for unroll in (1, 4):
        for gran in (32, 64):
            for tb in (128 << 20, 512 << 20, 1 << 30, 2 << 30, 4 << 30, 8 << 30, 16 << 30, 32 << 30, 64 << 30):
                if fg == 32 and tb not in (4 << 30, 64 << 30):
                    continue
                gb, gl = cs.probe_random_gather(0, tb, gran, 1 << 28, 2, unroll, fg)
                out.append(dict(l2_fetch=fg, unroll=unroll, granule=gran, table_mib=tb >> 20, gb_s=round(gb, 1), gloads_s=round(gl, 2)))
                print(out[-1], flush=True)
json.dump(out, open("gpurun_out/probe_sweep.json", "w"), indent=1)
