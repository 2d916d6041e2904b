#!/bin/bash
# pull ray cast v2: parity suite, memcheck of the smoke run, bench, per-kernel ncu durations + full capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pull3_pytest.txt 2>&1
tail -8 gpurun_out/r02_pull3_pytest.txt
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/r02_pull3_memcheck.txt 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/r02_pull3_memcheck.txt
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/r02_pull3_bench.json 2> gpurun_out/r02_pull3_bench.err
echo "bench rc=$?"; cat gpurun_out/r02_pull3_bench.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], d['e2e']['value'], d.get('kernel_ms_per_step'), d['counters_per_step'])"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'k_ray_' -s 610 -c 2 -o gpurun_out/r02_pull3_ray python scripts/step_times.py 256 305 308 > gpurun_out/r02_pull3_ray.log 2>&1
tail -3 gpurun_out/r02_pull3_ray.log
