#!/bin/bash
# round 2, call 6: wave balancing on one GPU's share of a 4- / 8-GPU day
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $OUT/r2c6_gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -3 $OUT/r2c6_gpu_tests.log
for sh in 8 4 2; do
timeout 900 python tools/probe.py --goals 1 --roots 0 --share $sh --days 3 --configs ";balance_waves=0;group_log=1;balance_waves=0,group_log=1" > $OUT/r2c6_probe_share$sh.log 2>&1
echo "probe share $sh exit $?"; cut -c1-700 $OUT/r2c6_probe_share$sh.log
done
