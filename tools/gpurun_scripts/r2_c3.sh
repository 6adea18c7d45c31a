#!/bin/bash
# round 2, call 3: median-based bucket width under adaptation, walk profile at the 10 M-trip shape, cfg4 rehearsal on one GPU
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -x -q -m gpu > $OUT/r2c3_gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -2 $OUT/r2c3_gpu_tests.log
timeout 900 python tools/adapt_probe.py --days 9 --every 2 > $OUT/r2c3_adapt_probe.log 2>&1
echo "adapt probe exit $?"; cut -c1-330 $OUT/r2c3_adapt_probe.log
B10="python bench.py --trips 10000000 --steps 1 --warmup 1 --no-secondary --no-cpu-baseline --e2e-steps 1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:walk_kernel --launch-skip 5 -c 1 -o $OUT/r2c3_walk10m -f $B10 > $OUT/r2c3_ncu_walk10m.log 2>&1
echo "ncu walk exit $?"; tail -3 $OUT/r2c3_ncu_walk10m.log | cut -c1-300
timeout 1500 python bench.py --workload cfg4_road5m --trips 1250000 --steps 2 --warmup 1 --construction-every 2 --per-day --no-secondary --no-cpu-baseline > $OUT/r2c3_bench_road5m_1gpu.json 2> $OUT/r2c3_bench_road5m_1gpu.err
echo "road5m exit $?"; tail -3 $OUT/r2c3_bench_road5m_1gpu.err
python -c "
import json; d=json.load(open('$OUT/r2c3_bench_road5m_1gpu.json')); print({k:d[k] for k in ('value','ms_per_step','load_hash','kernel_ms_serial_day','per_day_ms_rank0','road_construction','setup_seconds')}); print(d['roofline']['frac'], d['e2e']['value'])"
