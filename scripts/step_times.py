"""Per-scan kernel times and counters of PFSlam2D at the benchmark size (steady regime): where does the mean come from?"""
import sys
sys.path.insert(0, '.')
import numpy as np
from iris_lama_b200 import api, synth
P = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T0 = int(sys.argv[2]) if len(sys.argv) > 2 else 300
T1 = int(sys.argv[3]) if len(sys.argv) > 3 else 400
ds = synth.make_dataset("loop", T1, n_beams=1080)
g = api.PFSlam2D(api.PFSlam2D.Options(P, trans_thresh=0.05, rot_thresh=0.05, seed=42, timing=1))
g.setPrior(*ds.truth[0])
prev = None
rows = []
for t in range(T1):
    g.update(ds.scans[t], ds.odom[t])
    if t >= T0 - 1:
        g.getPose()
        ms, _ = g.kernelTimes()
        last, _ = g.counters()
        if prev is not None:
            rows.append((t, ms["match_ms"] - prev["match_ms"], ms["raycast_ms"] - prev["raycast_ms"], ms["brushfire_ms"] - prev["brushfire_ms"],
                         last["dm_pops"], last["ray_cells"], last["resampled"], last["detached"]))
        prev = ms
a = np.array(rows, float)
print("scan  match  raycast  brushfire  dm_pops ray_cells resampled detached")
for r in rows: print("%4d %6.3f %7.3f %8.3f %8d %9d %3d %6d" % r)
print("mean", a[:, 1:4].mean(0), "median", np.median(a[:, 1:4], 0), "max", a[:, 1:4].max(0))
