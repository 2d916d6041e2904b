#!/bin/bash
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $T scripts/sharded_native_check.py 16 30 360 0.02 2>&1 | tail -5 | tee gpurun_out/r02_shard2_check_small.txt
timeout 900 $T scripts/sharded_native_check.py 256 60 1080 0.0008 2>&1 | tail -5 | tee gpurun_out/r02_shard2_check_full.txt
timeout 900 $T bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r02_shard2_bench_native.json 2> gpurun_out/r02_shard2_bench_native.err; tail -2 gpurun_out/r02_shard2_bench_native.err
timeout 900 $T bench.py --gpus 2 --steps 100 --warmup 5 --sharded-impl python > gpurun_out/r02_shard2_bench_python.json 2> gpurun_out/r02_shard2_bench_python.err
python -c "
import json
for f in ('native','python'):
    try:
        d=json.loads(open('gpurun_out/r02_shard2_bench_%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['e2e']['value'], d['step_ms'])
    except Exception as e: print(f, 'ERR', e)
"
