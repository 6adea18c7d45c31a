#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "codes or plateau or ties" > $OUT/r2c20_gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -2 $OUT/r2c20_gpu_tests.log
timeout 900 python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline --trips 10000000 > $OUT/r2c20_bench10m.json 2> $OUT/r2c20_bench10m.err
python -c "
import json; d=json.load(open('$OUT/r2c20_bench10m.json')); print('10M codes(auto)', {k:d[k] for k in ('value','ms_per_step','load_hash','kernel_ms_serial_day')})"
B10="python bench.py --trips 10000000 --steps 1 --warmup 1 --no-secondary --no-cpu-baseline --e2e-steps 1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"walk_code_kernel|pred_code_kernel" --launch-skip 4 -c 2 -o $OUT/r2c20_walk_codes -f $B10 > $OUT/r2c20_ncu_walk_codes.log 2>&1
echo "ncu exit $?"; tail -2 $OUT/r2c20_ncu_walk_codes.log | cut -c1-200
