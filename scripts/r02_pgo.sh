#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_pgo.py -m gpu -x -q --durations=5 > gpurun_out/r02_pgo_pytest.txt 2>&1
tail -30 gpurun_out/r02_pgo_pytest.txt
python - <<'PY' 2>&1 | tee gpurun_out/r02_pgo_config5.txt
import sys, time
sys.path.insert(0, '.')
import numpy as np
from iris_lama_b200 import api, synth
truth, nodes, edges = synth.make_pose_graph(10000, 40001, seed=11, radius=3.0)
for rep in range(3):
    g = api.SimplePGO(nodes, edges)
    t = time.perf_counter(); ok = g.optimize(); dt = time.perf_counter() - t
    print("config5: 10000 poses,", len(edges) + 9999, "constraints: ok", ok, "status", g.status, g.report, "wall s", round(dt, 4))
PY
