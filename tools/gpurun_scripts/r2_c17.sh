#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1800 python -m pytest tests -x -q -m gpu > $OUT/r2c17_gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -3 $OUT/r2c17_gpu_tests.log
timeout 600 python tools/share_scan.py 8 > $OUT/r2c17_share_scan8.log 2>&1; cat $OUT/r2c17_share_scan8.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline > $OUT/r2c17_bench.json 2> $OUT/r2c17_bench.err
python -c "
import json; d=json.load(open('$OUT/r2c17_bench.json')); print({k:d[k] for k in ('value','ms_per_step','load_hash','kernel_ms_serial_day')}); print(d['roofline']['frac'], d['roofline']['frac_overlapped_region'], d['e2e']['value'], d['per_rank'])"
timeout 900 python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline --trips 10000000 > $OUT/r2c17_bench10m.json 2> $OUT/r2c17_bench10m.err
python -c "
import json; d=json.load(open('$OUT/r2c17_bench10m.json')); print('10M', {k:d[k] for k in ('value','ms_per_step','kernel_ms_serial_day')})"
timeout 900 python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline --trips 10000000 --options l2_persist=1 > $OUT/r2c17_bench10m_l2.json 2> $OUT/r2c17_bench10m_l2.err
python -c "
import json; d=json.load(open('$OUT/r2c17_bench10m_l2.json')); print('10M l2_persist', {k:d[k] for k in ('value','ms_per_step','kernel_ms_serial_day')})"
