# one GPU session: parity suite, both bench arms, ncu launch list + full capture of the three hot kernels (steady regime)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1240 -c 64 --csv --log-file gpurun_out/launches.csv python scripts/prof_run.py 256 330 > gpurun_out/prof1.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_match|k_raycast|k_brushfire" -s 930 -c 3 -o gpurun_out/final_kernels python scripts/prof_run.py 256 314 > gpurun_out/prof2.log 2>&1
cat gpurun_out/pytest_gpu.txt; tail -c 600 gpurun_out/bench_1gpu.json; tail -c 300 gpurun_out/bench_reference.json
