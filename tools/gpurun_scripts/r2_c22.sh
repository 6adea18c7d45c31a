#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
S4MPI_OPTIONS="walk_codes=1" timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_fullsize.py -x -q -m gpu > $OUT/r2c22_gpu_tests_codes.log 2>&1
echo "gpu tests (walk_codes=1 everywhere) exit $?"; tail -2 $OUT/r2c22_gpu_tests_codes.log
timeout 900 python bench.py --workload planar1m --steps 3 --warmup 2 --no-secondary --no-cpu-baseline > $OUT/r2c22_bench_planar1m.json 2> $OUT/r2c22_bench_planar1m.err
python -c "
import json; d=json.load(open('$OUT/r2c22_bench_planar1m.json')); print('planar1m', {k:d[k] for k in ('value','ms_per_step','load_hash','kernel_ms_serial_day','hops_per_trip')}, d['roofline']['frac'])"
timeout 900 python bench.py --steps 3 --warmup 2 --no-secondary --no-cpu-baseline --trips 10000000 > $OUT/r2c22_bench10m.json 2> $OUT/r2c22_bench10m.err
python -c "
import json; d=json.load(open('$OUT/r2c22_bench10m.json')); print('10M', {k:d[k] for k in ('value','ms_per_step','load_hash','kernel_ms_serial_day')})"
