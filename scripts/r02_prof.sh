#!/bin/bash
# ncu: launch list of 3 scans in the revisit regime + one full capture of the ray-cast kernels at scan 305
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_' -s 1830 -c 18 --csv --log-file gpurun_out/r02_prof_launches.csv python scripts/step_times.py 256 305 312 > gpurun_out/r02_prof_launches.log 2>&1
grep -o '"k_[a-z_]*[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","gpu__time_duration.sum","ns","[0-9]*"' gpurun_out/r02_prof_launches.csv | sed 's/(StoreView[^"]*//' | awk -F'","' '{print $1, $NF}' | head -30
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'k_ray_' -s 610 -c 2 -o gpurun_out/r02_prof_ray python scripts/step_times.py 256 305 308 > gpurun_out/r02_prof_ray.log 2>&1
ls -la gpurun_out/*.ncu-rep
