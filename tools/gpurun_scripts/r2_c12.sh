#!/bin/bash
# round 2, call 12: 31-tree groups as the default on cfg3-sized graphs: whole suite, bench with the secondary graphs, lag, shares
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1800 python -m pytest tests -x -q -m gpu > $OUT/r2c12_gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -3 $OUT/r2c12_gpu_tests.log
timeout 1500 python bench.py --steps 3 --warmup 3 > $OUT/r2c12_bench.json 2> $OUT/r2c12_bench.err
echo "bench exit $?"; tail -2 $OUT/r2c12_bench.err; python -c "
import json; d=json.load(open('$OUT/r2c12_bench.json')); print({k:d[k] for k in ('value','ms_per_step','load_hash','kernel_ms_serial_day')}); print(d['roofline']['frac'], d['e2e']['value']); s=d['secondary']
for k in s: print(k, {x:s[k][x] for x in ('ms_per_day','trips_per_sec','sssp_roofline_frac_rank0','sssp_ms_serial_day_rank0','walk_ms_serial_day_rank0')})"
timeout 600 python tools/probe.py --goals 1 --roots 0 --days 2 --configs "lag_factor=1.5;lag_factor=2;lag_factor=0.75;delta_factor=2" > $OUT/r2c12_probe_lag.log 2>&1
cut -c1-330 $OUT/r2c12_probe_lag.log
timeout 400 python tools/group_compare.py > $OUT/r2c12_group_compare.log 2>&1; grep -h "\"groups\"" $OUT/r2c12_group_compare.log | cut -c1-330
