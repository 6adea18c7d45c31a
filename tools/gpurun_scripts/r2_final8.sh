#!/bin/bash
# round 2, final multi-GPU lines: the default bench at N=8 (with the north-star day) and N=4
set -u
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29561 bench.py --gpus 8 --steps 5 --warmup 3 > $OUT/r2f8_bench_8gpu.json 2> $OUT/r2f8_bench_8gpu.err
echo "bench N=8 exit $?"
python -c "
import json; d=json.load(open('$OUT/r2f8_bench_8gpu.json')); print({k:d[k] for k in ('value','ms_per_step','load_hash','kernel_ms_serial_day')}); print(d['roofline']['frac'], d['e2e']['value']); s=d['secondary']
for k in s: print(k, {x:s[k][x] for x in ('ms_per_day','trips_per_sec','sssp_roofline_frac_rank0','sssp_ms_serial_day_rank0','walk_ms_serial_day_rank0','load_hash')})
for r in d['per_rank']: print(r)"
timeout 600 $TR --nproc-per-node 4 --master-port 29562 bench.py --gpus 4 --steps 5 --warmup 3 --no-secondary > $OUT/r2f8_bench_4gpu.json 2> $OUT/r2f8_bench_4gpu.err
echo "bench N=4 exit $?"; python -c "
import json; d=json.load(open('$OUT/r2f8_bench_4gpu.json')); print({k:d[k] for k in ('value','ms_per_step','load_hash','kernel_ms_serial_day')}); print(d['roofline']['frac'], d['e2e']['value'])
for r in d['per_rank']: print(r)"
