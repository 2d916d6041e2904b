#!/bin/bash
# one 8-GPU box: parity of the in-library sharded step at world 8 (small + full size, forced resampling => map migration), then the bench at N = 8 and 4
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
T8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611"
T4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29612"
nvidia-smi -L > gpurun_out/r02_scale8_gpus.txt; nproc >> gpurun_out/r02_scale8_gpus.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/r02_scale8_gpus.txt 2>/dev/null
timeout 300 $T8 scripts/sharded_native_check.py 64 40 360 0.02 2>&1 | tail -6 | tee gpurun_out/r02_scale8_check_small.txt
timeout 600 $T8 scripts/sharded_native_check.py 256 60 1080 0.0008 2>&1 | tail -6 | tee gpurun_out/r02_scale8_check_full.txt
timeout 600 $T8 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/r02_scale8_bench8.json 2> gpurun_out/r02_scale8_bench8.err; tail -3 gpurun_out/r02_scale8_bench8.err
timeout 600 $T4 bench.py --gpus 4 --steps 100 --warmup 5 > gpurun_out/r02_scale8_bench4.json 2> gpurun_out/r02_scale8_bench4.err; tail -3 gpurun_out/r02_scale8_bench4.err
timeout 600 python bench.py --gpus 1 --steps 100 --warmup 5 --no-cpu --no-regimes > gpurun_out/r02_scale8_bench1.json 2> gpurun_out/r02_scale8_bench1.err
python -c "
import json
for f in ('8','4','1'):
    try:
        d=json.loads(open('gpurun_out/r02_scale8_bench%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['e2e']['value'], d['step_ms'], d.get('weak_scaling'))
    except Exception as e: print(f, 'ERR', e)
"
