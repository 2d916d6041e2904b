#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_bf2_pytest.txt 2>&1
tail -5 gpurun_out/r02_bf2_pytest.txt
for v in 1 2; do
  echo "brushfire v$v revisit: $(LAMA_BRUSHFIRE=$v timeout 300 python scripts/step_times.py 256 300 340 | tail -1)"
  echo "brushfire v$v explore: $(LAMA_BRUSHFIRE=$v timeout 300 python scripts/explore_times.py 256 | tail -1)"
done 2>&1 | tee gpurun_out/r02_bf2_times.txt
LAMA_BRUSHFIRE=2 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
