#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pull4_pytest.txt 2>&1
tail -3 gpurun_out/r02_pull4_pytest.txt
for P in 32 256; do echo "P=$P pull: $(timeout 300 python scripts/step_times.py $P 300 340 | tail -1)"; done 2>&1 | tee gpurun_out/r02_pull4_times.txt
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'k_ray_' -s 610 -c 2 -o gpurun_out/r02_pull4_ray python scripts/step_times.py 256 305 308 > gpurun_out/r02_pull4_ray.log 2>&1
tail -1 gpurun_out/r02_pull4_ray.log
