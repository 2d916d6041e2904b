#!/bin/bash
# pull vs per-beam ray cast at the particle counts of an 8/4/2/1-GPU shard (kernel times of scans 300-340, mean match/raycast/brushfire ms)
mkdir -p gpurun_out
for P in 32 64 128 256; do
  for mode in pull walk; do
    if [ $mode = walk ]; then export LAMA_NO_PULL=1; else unset LAMA_NO_PULL; fi
    echo "P=$P $mode: $(timeout 300 python scripts/step_times.py $P 300 340 | tail -1)"
  done
done 2>&1 | tee gpurun_out/r02_cross.txt
