#!/bin/bash
# round 2, call 13: local-density clustering + join on shares, whole-wave launches, slim-ballot build A/B
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -x -q -m gpu > $OUT/r2c13_gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -2 $OUT/r2c13_gpu_tests.log
timeout 400 python tools/group_compare.py > $OUT/r2c13_group_compare.log 2>&1; grep -h "\"groups\"" $OUT/r2c13_group_compare.log | cut -c1-330
timeout 400 python tools/group_compare.py group_words=2 > $OUT/r2c13_group_compare_k15.log 2>&1; grep -h "\"groups\"" $OUT/r2c13_group_compare_k15.log | cut -c1-330
S4MPI_LIB=$PWD/streets4mpi_b200/libs4mpi_slim.so timeout 400 python tools/group_compare.py > $OUT/r2c13_group_compare_slim.log 2>&1; grep -h "\"groups\"" $OUT/r2c13_group_compare_slim.log | cut -c1-330
timeout 900 python bench.py --steps 3 --warmup 2 --no-secondary --no-cpu-baseline > $OUT/r2c13_bench.json 2> $OUT/r2c13_bench.err
python -c "
import json; d=json.load(open('$OUT/r2c13_bench.json')); print({k:d[k] for k in ('value','ms_per_step','load_hash','kernel_ms_serial_day')}); print(d['roofline']['frac'], d['roofline']['frac_overlapped_region'], d['e2e']['value'])"
S4MPI_LIB=$PWD/streets4mpi_b200/libs4mpi_slim.so timeout 900 python bench.py --steps 3 --warmup 2 --no-secondary --no-cpu-baseline > $OUT/r2c13_bench_slim.json 2> $OUT/r2c13_bench_slim.err
python -c "
import json; d=json.load(open('$OUT/r2c13_bench_slim.json')); print('slim', {k:d[k] for k in ('value','ms_per_step','load_hash','kernel_ms_serial_day')}); print(d['roofline']['frac'], d['roofline']['frac_overlapped_region'], d['e2e']['value'])"
