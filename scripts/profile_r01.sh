#!/bin/bash
# Round-1 evidence run (GPU box): host<->device copy bandwidth, full default bench, ncu launch list and one
# ncu --set full capture of the dominant kernel on the cfg2 index with 2M reads per launch.
mkdir -p gpurun_out
python - > gpurun_out/pcie.log 2>&1 <<'PY'
import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, (dst, src) in {"h2d": (d, h), "d2h": (h, d)}.items():
    best = 0
    for _ in range(4):
        torch.cuda.synchronize(); t = time.perf_counter(); dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
        best = max(best, n / (time.perf_counter() - t) / 1e9)
    print(name, "GB/s", round(best, 2))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t = time.perf_counter()
with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print("both directions at once: GB/s per direction", round(n / dt / 1e9, 2))
PY
cat gpurun_out/pcie.log
python bench.py > gpurun_out/bench_default.log 2>&1; tail -c 600 gpurun_out/bench_default.log
python bench.py --impl reference > gpurun_out/bench_reference.log 2>&1; tail -c 400 gpurun_out/bench_reference.log
CMD="python bench.py --reads 2000000 --steps 2 --warmup 1 --no-cpu --no-e2e"
$CMD > gpurun_out/plain_r01.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu_launches_r01.log 2>&1
$CMD > gpurun_out/plain_r01b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_seed_fast -s 1 -c 1 -f -o gpurun_out/prof_fast_r01 $CMD > gpurun_out/ncu_fast_r01.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r01.csv
