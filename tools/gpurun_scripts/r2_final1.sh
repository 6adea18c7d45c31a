#!/bin/bash
# round 2, final 1-GPU evidence: whole suite, the default bench line, launch list, full ncu captures of the final kernels
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1800 python -m pytest tests -x -q -m gpu > $OUT/r2f_gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -3 $OUT/r2f_gpu_tests.log
timeout 1500 python bench.py --steps 3 --warmup 3 > $OUT/r2f_bench.json 2> $OUT/r2f_bench.err
echo "bench exit $?"; tail -2 $OUT/r2f_bench.err | cut -c1-200
python -c "
import json; d=json.load(open('$OUT/r2f_bench.json')); print({k:d[k] for k in ('value','ms_per_step','load_hash','kernel_ms_serial_day')}); print(d['roofline']['frac'], d['roofline']['frac_overlapped_region'], d['e2e']['value'], d['cpu_baseline']); s=d['secondary']
for k in s: print(k, {x:s[k][x] for x in ('ms_per_day','trips_per_sec','sssp_roofline_frac_rank0','sssp_ms_serial_day_rank0','walk_ms_serial_day_rank0')})"
B="python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline --e2e-steps 1"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r2f_launches.csv $B > $OUT/r2f_ncu_launches.log 2>&1
echo "launch list exit $?"
P="python tools/probe.py --goals 1 --roots 0 --days 1"
$P > $OUT/r2f_probe_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:sssp_group -c 1 -o $OUT/r2f_group -f $P > $OUT/r2f_ncu_group.log 2>&1
echo "ncu group exit $?"; tail -1 $OUT/r2f_probe_plain.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:walk_kernel --launch-skip 3 -c 1 -o $OUT/r2f_walk -f $B > $OUT/r2f_ncu_walk.log 2>&1
echo "ncu walk exit $?"
timeout 900 python bench.py --steps 3 --warmup 2 --no-secondary --no-cpu-baseline --trips 10000000 > $OUT/r2f_bench10m.json 2> $OUT/r2f_bench10m.err
python -c "
import json; d=json.load(open('$OUT/r2f_bench10m.json')); print('10M', {k:d[k] for k in ('value','ms_per_step','load_hash','kernel_ms_serial_day')})"
