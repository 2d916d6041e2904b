#!/bin/bash
mkdir -p gpurun_out
LAMA_BRUSHFIRE=2 timeout 900 ncu --set full --import-source on --clock-control none -k regex:'k_brushfire2' -s 20 -c 1 -o gpurun_out/r02_bf2 python scripts/explore_times.py 256 > gpurun_out/r02_bf2_prof.log 2>&1
tail -2 gpurun_out/r02_bf2_prof.log
