"""kernel times in the exploration regime (scans 6..45 of a fresh map) -- mean match / raycast / brushfire ms"""
import sys
sys.path.insert(0, '.')
import numpy as np
from iris_lama_b200 import api, synth
P = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ds = synth.make_dataset("loop", 46, n_beams=1080)
g = api.PFSlam2D(api.PFSlam2D.Options(P, trans_thresh=0.05, rot_thresh=0.05, seed=42, timing=1))
g.setPrior(*ds.truth[0])
prev = None; rows = []
for t in range(46):
    g.update(ds.scans[t], ds.odom[t])
    if t >= 5:
        g.getPose()
        ms, _ = g.kernelTimes()
        last, _ = g.counters()
        if prev is not None: rows.append((ms["match_ms"] - prev["match_ms"], ms["raycast_ms"] - prev["raycast_ms"], ms["brushfire_ms"] - prev["brushfire_ms"], last["dm_pops"]))
        prev = ms
a = np.array(rows)
print("explore mean", a[:, :3].mean(0), "median", np.median(a[:, :3], 0), "max", a[:, :3].max(0), "pops/scan", a[:, 3].mean())
