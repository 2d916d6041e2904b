#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_lc_pytest.txt 2>&1
tail -15 gpurun_out/r02_lc_pytest.txt
for P in 256; do echo "P=$P: $(timeout 300 python scripts/step_times.py $P 300 340 | tail -1)"; done 2>&1 | tee gpurun_out/r02_lc_times.txt
