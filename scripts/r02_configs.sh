#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_configs.py -q -m gpu -s 2>&1 | tail -12 > gpurun_out/r02_configs_pytest.txt
cat gpurun_out/r02_configs_pytest.txt
timeout 900 python scripts/config4_parity.py 1501 16 0.001 > gpurun_out/r02_config4_parity_forced_resampling_1500.txt 2>&1; tail -3 gpurun_out/r02_config4_parity_forced_resampling_1500.txt
timeout 1500 python scripts/config4_parity.py 5001 16 > gpurun_out/r02_config4_parity_5000.txt 2>&1; tail -3 gpurun_out/r02_config4_parity_5000.txt
