#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/bf_phases.py 330 2>&1 | grep "^bf" > gpurun_out/r02_bf_phases_330.txt
wc -l gpurun_out/r02_bf_phases_*.txt
