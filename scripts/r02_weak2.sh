#!/bin/bash
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $T bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r02_weak2_bench.json 2> gpurun_out/r02_weak2_bench.err; tail -3 gpurun_out/r02_weak2_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_weak2_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d.get('weak_scaling'))
"
