#!/bin/bash
# round 2, 8-GPU call: the default bench line at N=8 (with the north-star day) and BASELINE configs[3] as specified
set -u
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 > $OUT/r2c8_bench_8gpu.json 2> $OUT/r2c8_bench_8gpu.err
echo "bench N=8 exit $?"; tail -2 $OUT/r2c8_bench_8gpu.err | cut -c1-300
python -c "
import json; d=json.load(open('$OUT/r2c8_bench_8gpu.json')); print({k:d[k] for k in ('value','ms_per_step','load_hash','kernel_ms_serial_day')}); print(d['roofline']['frac'], d['e2e']['value']); print(json.dumps(d['secondary'])[:1500])"
timeout 900 $TR --master-port 29532 bench.py --gpus 8 --workload cfg4_road5m --steps 30 --warmup 1 --construction-every 5 --per-day --no-secondary > $OUT/r2c8_bench_cfg4_8gpu.json 2> $OUT/r2c8_bench_cfg4_8gpu.err
echo "cfg4 N=8 exit $?"; tail -3 $OUT/r2c8_bench_cfg4_8gpu.err | cut -c1-300
python -c "
import json; d=json.load(open('$OUT/r2c8_bench_cfg4_8gpu.json')); print({k:d[k] for k in ('value','ms_per_step','load_hash','kernel_ms_serial_day','per_day_ms_rank0','road_construction','setup_seconds')}); print(d['roofline']['frac'], d['e2e'])"
