#!/bin/bash
# Round-1 final evidence (one B200): default bench, reference arm, launch list, ncu capture of the dominant kernel, config sweep.
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_default.log 2>&1; tail -c 300 gpurun_out/bench_default.log; echo
python bench.py --impl reference > gpurun_out/bench_reference.log 2>&1; tail -c 200 gpurun_out/bench_reference.log; echo
python bench.py --probe --no-cpu --no-e2e --steps 2 --warmup 3 > gpurun_out/bench_probe.log 2>&1
timeout 900 python scripts/config_sweep.py > gpurun_out/config_sweep.json 2> gpurun_out/config_sweep.log; tail -3 gpurun_out/config_sweep.log
CMD="python bench.py --reads 2000000 --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/plain_final.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_seed|k_collect|k_mem_counts|k_sa_resolve|k_pack|DeviceScan' -s 30 -c 60 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_launches_final.log 2>&1
$CMD > gpurun_out/plain_final2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_seed_fast -s 3 -c 1 -f -o gpurun_out/prof_fast_final $CMD > gpurun_out/ncu_fast_final.log 2>&1
ls -la gpurun_out/*final*
