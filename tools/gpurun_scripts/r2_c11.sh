#!/bin/bash
# round 2, call 11: 31 trees per group (4 words per lane)
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $OUT/r2c11_gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -3 $OUT/r2c11_gpu_tests.log
timeout 900 python tools/probe.py --goals 1 --roots 0 --days 2 --configs "group_log=1;group_log=1,group_words=4;group_log=1,group_words=4,block_threads=512;group_log=1,group_words=4,lag_factor=0.5;group_log=1,group_words=4,block_threads=512,lag_factor=0.5;group_log=1,group_words=4,delta_factor=1.0" > $OUT/r2c11_probe_day.log 2>&1
echo "probe day exit $?"; cut -c1-900 $OUT/r2c11_probe_day.log
