#!/bin/bash
# round 2, 8-GPU call 2: the default bench line at N=8 and N=4 after wave packing / 31-tree groups
set -u
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29551 bench.py --gpus 8 --steps 5 --warmup 3 > $OUT/r2c14_bench_8gpu.json 2> $OUT/r2c14_bench_8gpu.err
echo "bench N=8 exit $?"; tail -2 $OUT/r2c14_bench_8gpu.err | cut -c1-200
python -c "
import json; d=json.load(open('$OUT/r2c14_bench_8gpu.json')); print({k:d[k] for k in ('value','ms_per_step','load_hash','kernel_ms_serial_day')}); print(d['roofline']['frac'], d['e2e']['value']); s=d['secondary']
for k in s: print(k, {x:s[k][x] for x in ('ms_per_day','trips_per_sec','sssp_roofline_frac_rank0','sssp_ms_serial_day_rank0','walk_ms_serial_day_rank0')})"
timeout 600 $TR --nproc-per-node 4 --master-port 29552 bench.py --gpus 4 --steps 5 --warmup 3 --no-secondary > $OUT/r2c14_bench_4gpu.json 2> $OUT/r2c14_bench_4gpu.err
echo "bench N=4 exit $?"; python -c "
import json; d=json.load(open('$OUT/r2c14_bench_4gpu.json')); print({k:d[k] for k in ('value','ms_per_step','load_hash','kernel_ms_serial_day')}); print(d['roofline']['frac'], d['e2e']['value'])"
