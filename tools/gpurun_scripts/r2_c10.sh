#!/bin/bash
# round 2, 2-GPU call: multi-GPU parity tests, cfg5 weak-scaling pair, the default bench line at N=2
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > $OUT/r2c10_gpu_multi_tests.log 2>&1
echo "multi tests exit $?"; tail -3 $OUT/r2c10_gpu_multi_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29541 tools/run_cfg5_sample.py --trees 420 --no-oracle > $OUT/r2c10_cfg5_weak_2gpu.json 2> $OUT/r2c10_cfg5_weak_2gpu.err
echo "cfg5 2-GPU exit $?"; tail -2 $OUT/r2c10_cfg5_weak_2gpu.err | cut -c1-300; cut -c1-2500 $OUT/r2c10_cfg5_weak_2gpu.json
timeout 600 $TR --master-port 29542 bench.py --gpus 2 --steps 3 --warmup 3 > $OUT/r2c10_bench_2gpu.json 2> $OUT/r2c10_bench_2gpu.err
echo "bench N=2 exit $?"; python -c "
import json; d=json.load(open('$OUT/r2c10_bench_2gpu.json')); print({k:d[k] for k in ('value','ms_per_step','load_hash','kernel_ms_serial_day')}); print(d['roofline']['frac'], d['e2e']['value']); print(json.dumps(d['secondary'])[:600])"
