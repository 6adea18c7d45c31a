#!/bin/bash
# round 2, call 2: suite with the new tests, adaptation diagnostics on the road-like graph, walk check at 10 M trips
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1800 python -m pytest tests -x -q -m gpu > $OUT/r2c2_gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -3 $OUT/r2c2_gpu_tests.log
timeout 900 python tools/adapt_probe.py --days 9 --every 2 > $OUT/r2c2_adapt_probe.log 2>&1
echo "adapt probe exit $?"; cat $OUT/r2c2_adapt_probe.log | cut -c1-700
timeout 900 python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline --trips 10000000 > $OUT/r2c2_bench10m.json 2> $OUT/r2c2_bench10m.err
echo "bench10m exit $?"
python -c "
import json; d=json.load(open('$OUT/r2c2_bench10m.json')); print('10M', {k:d[k] for k in ('value','ms_per_step','load_hash','kernel_ms_serial_day')})"
