#!/bin/bash
# Round-2 GPU call AD: k_chain_build with the sequence hint -- chain parity tests, kernel times.
set -u
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_gpu_chain.py -m gpu -x -q ) > $OUT/ad_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/ad_pytest.log; tail -4 $OUT/ad_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_chain' --csv --log-file $OUT/ad_chain_launches.csv python scripts/r02_chain_time.py 2000000 > $OUT/ad_ncu1.log 2>&1; echo "ncu times rc=$?"
grep k_chain $OUT/ad_chain_launches.csv | awk -F'","' '{print substr($5,1,20), $NF}' | tail -3
