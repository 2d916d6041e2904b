#!/bin/bash
# round-2 first call: sanity (pytest -m gpu), lane utilisation of the round-1 ray cast, CPU-arm facts of this host
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_probe_gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_probe_pytest.txt 2>&1
tail -3 gpurun_out/r02_probe_pytest.txt
timeout 600 ncu --metrics smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,l1tex__t_requests_pipe_lsu_mem_global_op_red.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum \
  --clock-control none -k regex:'k_raycast|k_brushfire|k_match' -s 915 -c 6 --csv --log-file gpurun_out/r02_probe_lanes.csv python scripts/step_times.py 256 305 308 > gpurun_out/r02_probe_lanes.log 2>&1
tail -8 gpurun_out/r02_probe_lanes.csv
timeout 900 python scripts/cpu_probe.py 120 8 16 32 64 128 > gpurun_out/r02_probe_cpu.txt 2>&1
cat gpurun_out/r02_probe_cpu.txt
