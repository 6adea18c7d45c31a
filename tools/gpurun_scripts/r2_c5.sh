#!/bin/bash
# round 2, call 5: clusters on cfg3 (whole day and an 8-GPU share of it), split unrolling under adaptation
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $OUT/r2c5_gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -3 $OUT/r2c5_gpu_tests.log
timeout 900 python tools/probe.py --goals 1 --roots 0 --days 2 --configs ";cluster=2;cluster=4;cluster=8;lanes=1;lanes=1,cluster=2" > $OUT/r2c5_probe_day.log 2>&1
echo "probe day exit $?"; cat $OUT/r2c5_probe_day.log
timeout 900 python tools/probe.py --goals 1 --roots 2622 --days 3 --configs ";cluster=2;cluster=4;cluster=8;lanes=1,cluster=2;lanes=1,cluster=4;lanes=1,cluster=8" > $OUT/r2c5_probe_share8.log 2>&1
echo "probe share exit $?"; cat $OUT/r2c5_probe_share8.log
timeout 900 python tools/adapt_probe.py --days 9 --every 2 > $OUT/r2c5_adapt_probe.log 2>&1
echo "adapt probe exit $?"; cut -c1-250 $OUT/r2c5_adapt_probe.log
