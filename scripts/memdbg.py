import sys; sys.path.insert(0,'.')
import numpy as np
from iris_lama_b200 import api, synth
from oracle import pyoracle as po
P, T = 8, 4
ds = synth.make_dataset("room", T, n_beams=180)
g = api.PFSlam2D(api.PFSlam2D.Options(P, trans_thresh=0.05, rot_thresh=0.05, seed=5))
o = po.PFSlam2D(po.PFOptions.defaults(P, trans_thresh=0.05, rot_thresh=0.05, seed=5))
g.setPrior(*ds.truth[0]); o.set_prior(*ds.truth[0])
for t in range(T):
    g.update(ds.scans[t], ds.odom[t]); o.update(ds.scans[t], ds.odom[t])
    print(t, "mem", g.getMemoryUsage(), o.memory_usage())
    for p in (0, 3):
        print("  p", p, "occ", g.mapBounds(p, 0), o.occ_bounds(p), "dm", g.mapBounds(p, 1), o.dm_bounds(p))
