#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python tools/probe.py --goals 1 --roots 0 --days 2 --configs "group_log=1;group_log=1,lag_factor=0.5;group_log=1,lag_factor=1;group_log=1,lag_factor=2;group_log=1,delta_factor=1.5;group_log=1,delta_factor=6;group_log=1,group_words=1;group_log=1,block_threads=512;group_log=1,early_advance=0;group_log=1,early_advance=8;group_log=1,adaptive_delta=0;group_log=1,min_batches=16;group_lanes=0" > $OUT/r2p2_probe_grid.log 2>&1
echo "probe grid exit $?"
timeout 900 python tools/probe.py --planar 500000 --goals 1 --roots 0 --days 2 --configs "group_log=1;group_log=1,lag_factor=1;group_log=1,block_threads=512;group_log=1,block_threads=1024;group_log=1,group_words=1;group_lanes=0" > $OUT/r2p2_probe_planar.log 2>&1
echo "probe planar exit $?"
cat $OUT/r2p2_probe_grid.log $OUT/r2p2_probe_planar.log
