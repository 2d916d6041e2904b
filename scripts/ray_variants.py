"""Developer experiment (LAMA_PHASE_TIMING build): k_raycast time with the RED / the LDS removed, after 300 correct scans."""
import os, sys
sys.path.insert(0, '.')
import numpy as np
from iris_lama_b200 import api, synth
ds = synth.make_dataset("loop", 400, n_beams=1080)
g = api.PFSlam2D(api.PFSlam2D.Options(256, trans_thresh=0.05, rot_thresh=0.05, seed=42, timing=1))
g.setPrior(*ds.truth[0])
def run(t0, t1):
    g.getPose(); a, _ = g.kernelTimes()
    for t in range(t0, t1): g.update(ds.scans[t], ds.odom[t])
    g.getPose(); b, _ = g.kernelTimes()
    return {k: (b[k] - a[k]) / (t1 - t0) for k in a}
run(0, 300)
t = 300
for dbg in (0, 1, 3, 0):
    os.environ["LAMA_RAY_DEBUG"] = str(dbg)
    print("debug", dbg, run(t, t + 20), flush=True)
    t += 20
