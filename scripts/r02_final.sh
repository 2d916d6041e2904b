#!/bin/bash
# one GPU session at the end of round 2: parity suite, both bench arms, ncu launch list + full capture of the hot kernels (revisit regime)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/r02_final_pytest_gpu.txt
timeout 1500 python bench.py > gpurun_out/r02_final_bench_1gpu.json 2> gpurun_out/r02_final_bench_1gpu.err
timeout 900 python bench.py --impl reference > gpurun_out/r02_final_bench_reference_arm.json 2> gpurun_out/r02_final_bench_reference_arm.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1240 -c 64 --csv --log-file gpurun_out/r02_final_launches.csv python scripts/prof_run.py 256 330 > gpurun_out/prof1.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_match|k_raycast|k_brushfire" -s 930 -c 3 -o gpurun_out/r02_final_kernels python scripts/prof_run.py 256 314 > gpurun_out/prof2.log 2>&1
LAMA_PULL_MAX_PARTICLES=48 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_ray_pull|k_ray_setup" -s 620 -c 2 -o gpurun_out/r02_final_pull32 python scripts/prof_run.py 32 314 > gpurun_out/prof3.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_final_smoke.txt 2>&1
cat gpurun_out/r02_final_pytest_gpu.txt; tail -2 gpurun_out/r02_final_smoke.txt; tail -c 400 gpurun_out/r02_final_bench_1gpu.err; tail -c 300 gpurun_out/r02_final_bench_reference_arm.json
