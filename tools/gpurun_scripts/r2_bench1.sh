#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/r2b1_gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -3 $OUT/r2b1_gpu_tests.log
timeout 1500 python bench.py --steps 3 --warmup 3 > $OUT/r2b1_bench.json 2> $OUT/r2b1_bench.err
echo "bench exit $?"; tail -5 $OUT/r2b1_bench.err; cat $OUT/r2b1_bench.json
