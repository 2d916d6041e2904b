#!/bin/bash
# 8-GPU box after the rng_skip change: parity at world 8 again (small + full size), then the N = 8 bench (strong + weak)
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
T8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611"
timeout 300 $T8 scripts/sharded_native_check.py 64 40 360 0.02 2>&1 | tail -2 | tee gpurun_out/r02_scale8b_check_small.txt
timeout 600 $T8 scripts/sharded_native_check.py 256 60 1080 0.0008 2>&1 | tail -2 | tee gpurun_out/r02_scale8b_check_full.txt
timeout 600 $T8 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/r02_scale8b_bench8.json 2> gpurun_out/r02_scale8b_bench8.err; tail -3 gpurun_out/r02_scale8b_bench8.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_scale8b_bench8.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['step_ms'], d.get('weak_scaling'))
"
