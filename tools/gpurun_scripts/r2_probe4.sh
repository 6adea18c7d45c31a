#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/r2p4_gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -4 $OUT/r2p4_gpu_tests.log
timeout 900 python tools/probe.py --goals 1 --roots 0 --days 2 --configs "group_log=1;group_log=1,lag_factor=0.5;group_log=1,lag_factor=2;group_log=1,delta_factor=1;group_log=1,delta_factor=2;group_log=1,block_threads=512;group_log=1,early_advance=4;group_log=1,group_words=1" > $OUT/r2p4_probe_grid.log 2>&1
echo "probe grid exit $?"
timeout 900 python tools/probe.py --goals 1 --roots 0 --days 2 --rank 0 --configs "group_log=1" > $OUT/r2p4_probe_grid_norank.log 2>&1
timeout 900 python tools/probe.py --planar 500000 --goals 1 --roots 0 --days 2 --configs "group_log=1;group_log=1,block_threads=128;group_log=1,block_threads=512;group_log=1,block_threads=1024;group_log=1,lag_factor=2;group_log=1,delta_factor=3;group_lanes=0" > $OUT/r2p4_probe_planar.log 2>&1
echo "probe planar exit $?"
timeout 900 python tools/probe.py --planar 500000 --goals 1 --roots 0 --days 2 --rank 0 --configs "group_log=1" > $OUT/r2p4_probe_planar_norank.log 2>&1
cat $OUT/r2p4_probe_grid.log $OUT/r2p4_probe_grid_norank.log $OUT/r2p4_probe_planar.log $OUT/r2p4_probe_planar_norank.log
