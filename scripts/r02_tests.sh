#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r02_tests_pytest.txt 2>&1
tail -25 gpurun_out/r02_tests_pytest.txt
for P in 32 256; do echo "P=$P forced pull: $(LAMA_PULL_MAX_PARTICLES=100000 timeout 300 python scripts/step_times.py $P 300 340 | tail -1)"; done 2>&1 | tee gpurun_out/r02_tests_times.txt
