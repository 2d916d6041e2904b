#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_bench1_pytest.txt 2>&1
tail -5 gpurun_out/r02_bench1_pytest.txt
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_bench1.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-regimes > gpurun_out/r02_bench1_k20.json 2> gpurun_out/r02_bench1_k20.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench1_ref.json 2> gpurun_out/r02_bench1_ref.err
