#!/bin/bash
# Round-2 GPU call AB: the whole GPU suite and smoke() on the final build.
set -u
OUT=gpurun_out; mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $OUT/ab_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/ab_pytest.log; tail -6 $OUT/ab_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/ab_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/ab_smoke.log
