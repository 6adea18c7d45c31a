#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -x -q -m gpu > $OUT/r2p3_parity_default.log 2>&1
echo "parity default exit $?"; tail -3 $OUT/r2p3_parity_default.log
for v in "group_words=1" "group_lanes=4" "lag_factor=1"; do
    tag=$(echo "$v" | tr '=,' '__')
    S4MPI_OPTIONS="$v" timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $OUT/r2p3_parity_$tag.log 2>&1
    echo "parity $v exit $?"; tail -1 $OUT/r2p3_parity_$tag.log
done
timeout 900 python tools/probe.py --goals 1 --roots 0 --days 2 --configs "group_log=1;group_log=1,delta_factor=2;group_log=1,delta_factor=1.5;group_log=1,delta_factor=1;group_log=1,delta_factor=0.7;group_log=1,delta_factor=1.5,lag_factor=1;group_log=1,delta_factor=1.5,early_advance=0;group_log=1,delta_factor=1.5,adaptive_delta=0;group_log=1,delta_factor=1.5,block_threads=512;group_log=1,delta_factor=1.5,group_words=1" > $OUT/r2p3_probe_grid.log 2>&1
echo "probe grid exit $?"
timeout 900 python tools/probe.py --planar 500000 --goals 1 --roots 0 --days 2 --configs "group_log=1;group_log=1,lag_factor=1;group_log=1,block_threads=512;group_log=1,block_threads=1024;group_log=1,delta_factor=1.5,block_threads=1024;group_log=1,delta_factor=6,block_threads=1024" > $OUT/r2p3_probe_planar.log 2>&1
echo "probe planar exit $?"
cat $OUT/r2p3_probe_grid.log $OUT/r2p3_probe_planar.log
