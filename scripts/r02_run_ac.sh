#!/bin/bash
# Round-2 GPU call AC: the chaining kernels on cfg2 (2 M reads): times, and a full ncu capture of k_chain_build.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python scripts/r02_chain_time.py 2000000 > $OUT/ac_chain_time.log 2>&1; echo "chain time rc=$?"; cat $OUT/ac_chain_time.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_chain' --csv --log-file $OUT/ac_chain_launches.csv python scripts/r02_chain_time.py 2000000 > $OUT/ac_ncu1.log 2>&1; echo "ncu times rc=$?"
grep k_chain $OUT/ac_chain_launches.csv | awk -F'","' '{print $5, $NF}' | tail -9
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_chain_build -s 2 -c 1 -o $OUT/ac_prof_chain -f python scripts/r02_chain_time.py 2000000 > $OUT/ac_ncu2.log 2>&1; echo "ncu full rc=$?"
