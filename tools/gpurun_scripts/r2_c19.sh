#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $OUT/r2c19_gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -3 $OUT/r2c19_gpu_tests.log
S4MPI_OPTIONS="walk_codes=1" timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_fullsize.py -x -q -m gpu > $OUT/r2c19_gpu_tests_codes.log 2>&1
echo "gpu tests (walk_codes=1 everywhere) exit $?"; tail -3 $OUT/r2c19_gpu_tests_codes.log
timeout 900 python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline --trips 10000000 > $OUT/r2c19_bench10m.json 2> $OUT/r2c19_bench10m.err
python -c "
import json; d=json.load(open('$OUT/r2c19_bench10m.json')); print('10M codes(auto)', {k:d[k] for k in ('value','ms_per_step','load_hash','kernel_ms_serial_day')})"
timeout 900 python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline --trips 10000000 --options walk_codes=0 > $OUT/r2c19_bench10m_nocodes.json 2> $OUT/r2c19_bench10m_nocodes.err
python -c "
import json; d=json.load(open('$OUT/r2c19_bench10m_nocodes.json')); print('10M no codes', {k:d[k] for k in ('value','ms_per_step','load_hash','kernel_ms_serial_day')})"
timeout 900 python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline --options walk_codes=1 > $OUT/r2c19_bench1m_codes.json 2> $OUT/r2c19_bench1m_codes.err
python -c "
import json; d=json.load(open('$OUT/r2c19_bench1m_codes.json')); print('1M codes forced', {k:d[k] for k in ('value','ms_per_step','load_hash','kernel_ms_serial_day')})"
