#!/bin/bash
# round 2, call 4: cluster-shared groups (parity + cfg5 shape A/B), regression bench
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $OUT/r2c4_gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -4 $OUT/r2c4_gpu_tests.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-secondary --no-cpu-baseline > $OUT/r2c4_bench.json 2> $OUT/r2c4_bench.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('$OUT/r2c4_bench.json')); print({k:d[k] for k in ('value','ms_per_step','load_hash','kernel_ms_serial_day')}); print(d['roofline']['frac'], d['e2e']['value'])"
timeout 1200 python tools/run_cfg5_sample.py --trees 420 > $OUT/r2c4_cfg5_cluster_auto.json 2> $OUT/r2c4_cfg5_cluster_auto.err
echo "cfg5 auto exit $?"; tail -2 $OUT/r2c4_cfg5_cluster_auto.err; cat $OUT/r2c4_cfg5_cluster_auto.json
timeout 1200 python tools/run_cfg5_sample.py --trees 420 --no-oracle --options cluster=1 > $OUT/r2c4_cfg5_cluster_1.json 2> $OUT/r2c4_cfg5_cluster_1.err
echo "cfg5 cluster=1 exit $?"; tail -2 $OUT/r2c4_cfg5_cluster_1.err; cat $OUT/r2c4_cfg5_cluster_1.json
