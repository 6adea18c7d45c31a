#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -x -q -m gpu > $OUT/r2c21_gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -2 $OUT/r2c21_gpu_tests.log
for w in planar1m road1m; do
timeout 900 python bench.py --workload $w --steps 3 --warmup 2 --no-secondary --no-cpu-baseline > $OUT/r2c21_bench_$w.json 2> $OUT/r2c21_bench_$w.err
python -c "
import json; d=json.load(open('$OUT/r2c21_bench_$w.json')); print('$w', {k:d[k] for k in ('value','ms_per_step','load_hash','kernel_ms_serial_day','hops_per_trip')}, d['roofline']['frac'])"
done
timeout 900 python bench.py --workload planar1m --steps 3 --warmup 2 --no-secondary --no-cpu-baseline --options walk_codes=0 > $OUT/r2c21_bench_planar1m_nocodes.json 2> $OUT/r2c21_bench_planar1m_nocodes.err
python -c "
import json; d=json.load(open('$OUT/r2c21_bench_planar1m_nocodes.json')); print('planar1m walk_codes=0', {k:d[k] for k in ('value','ms_per_step','load_hash','kernel_ms_serial_day')})"
timeout 900 python bench.py --steps 3 --warmup 2 --no-secondary --no-cpu-baseline --trips 10000000 > $OUT/r2c21_bench10m.json 2> $OUT/r2c21_bench10m.err
python -c "
import json; d=json.load(open('$OUT/r2c21_bench10m.json')); print('10M', {k:d[k] for k in ('value','ms_per_step','load_hash','kernel_ms_serial_day')})"
