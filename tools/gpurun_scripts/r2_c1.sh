#!/bin/bash
# round 2, call 1: full GPU suite, default bench line, cfg4 rehearsal on one GPU, launch list + full ncu captures
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/r2c1_gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -3 $OUT/r2c1_gpu_tests.log
timeout 1500 python bench.py --steps 3 --warmup 3 > $OUT/r2c1_bench.json 2> $OUT/r2c1_bench.err
echo "bench exit $?"; tail -3 $OUT/r2c1_bench.err
timeout 900 python bench.py --workload road1m --steps 6 --warmup 1 --construction-every 2 --no-secondary --no-cpu-baseline > $OUT/r2c1_bench_road1m_adapt.json 2> $OUT/r2c1_bench_road1m_adapt.err
echo "road1m exit $?"; tail -2 $OUT/r2c1_bench_road1m_adapt.err
B="python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline --e2e-steps 1"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r2c1_launches.csv $B > $OUT/r2c1_ncu_launches.log 2>&1
echo "launch list exit $?"
P="python tools/probe.py --goals 1 --roots 8880 --days 1"
$P > $OUT/r2c1_probe_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:sssp_group -c 1 -o $OUT/r2c1_group -f $P > $OUT/r2c1_ncu_group.log 2>&1
echo "ncu group exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -c 1 -o $OUT/r2c1_walk -f $P > $OUT/r2c1_ncu_walk.log 2>&1
echo "ncu walk exit $?"; tail -2 $OUT/r2c1_probe_plain.log
python -c "
import json; d=json.load(open('$OUT/r2c1_bench.json')); print({k:d[k] for k in ('value','ms_per_step','load_hash','kernel_ms_serial_day')}); print(d['roofline']['frac'], d['e2e']['value'])
d=json.load(open('$OUT/r2c1_bench_road1m_adapt.json')); print({k:d[k] for k in ('value','ms_per_step','load_hash','kernel_ms_serial_day')}); print(d['roofline']['frac'])"
