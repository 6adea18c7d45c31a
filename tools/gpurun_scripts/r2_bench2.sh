#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/r2b2_gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -3 $OUT/r2b2_gpu_tests.log
S4MPI_OPTIONS="fused_walk=0" timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -x -q -m gpu > $OUT/r2b2_parity_unfused.log 2>&1
echo "unfused parity exit $?"; tail -1 $OUT/r2b2_parity_unfused.log
timeout 1500 python bench.py --steps 3 --warmup 3 --extra-graphs "" --no-cpu-baseline > $OUT/r2b2_bench.json 2> $OUT/r2b2_bench.err
echo "bench exit $?"; tail -3 $OUT/r2b2_bench.err; python -c "
import json; d=json.load(open('$OUT/r2b2_bench.json')); print({k:d[k] for k in ('value','ms_per_step','load_hash','kernel_ms_serial_day')}); print(d['roofline']['frac'], d['e2e']['value'], d['secondary'])"
timeout 900 python bench.py --steps 3 --warmup 2 --extra-graphs "" --no-cpu-baseline --no-secondary --trips 10000000 > $OUT/r2b2_bench10m.json 2> $OUT/r2b2_bench10m.err
python -c "
import json; d=json.load(open('$OUT/r2b2_bench10m.json')); print('10M', {k:d[k] for k in ('value','ms_per_step','load_hash','kernel_ms_serial_day')})"
