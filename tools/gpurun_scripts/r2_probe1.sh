#!/bin/bash
# round 2, first GPU run of the source-batched SSSP kernel: parity, then timings against the one-tree-per-CTA kernel
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/r2p1_smi.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -x -q -m gpu > $OUT/r2p1_parity_default.log 2>&1
echo "parity default exit $?" | tee -a $OUT/r2p1.log
for v in "group_words=1" "group_lanes=4" "group_lanes=4,group_words=1" "lag_factor=1" "group_lanes=0"; do
    tag=$(echo "$v" | tr '=,' '__')
    S4MPI_OPTIONS="$v" timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $OUT/r2p1_parity_$tag.log 2>&1
    echo "parity $v exit $?" | tee -a $OUT/r2p1.log
done
timeout 900 python tools/probe.py --goals 1 --roots 0 --days 2 --configs ";group_words=1;group_lanes=4;group_lanes=4,group_words=1;group_lanes=0;block_threads=512;block_threads=256;block_threads=512,group_words=1;lag_factor=1;lag_factor=0.5,block_threads=512;delta_factor=6;delta_factor=1.5" > $OUT/r2p1_probe_grid.log 2>&1
echo "probe grid exit $?" | tee -a $OUT/r2p1.log
timeout 900 python tools/probe.py --planar 500000 --goals 1 --roots 0 --days 2 --configs ";group_words=1;group_lanes=4;group_lanes=0;block_threads=256;block_threads=128;block_threads=64;lag_factor=1;lag_factor=1,group_words=1" > $OUT/r2p1_probe_planar.log 2>&1
echo "probe planar exit $?" | tee -a $OUT/r2p1.log
tail -5 $OUT/r2p1_parity_default.log
cat $OUT/r2p1_probe_grid.log $OUT/r2p1_probe_planar.log
