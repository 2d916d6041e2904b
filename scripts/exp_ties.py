"""Does heap tie order change the distance field?  Slam2D literal vs shuffled tie order."""
import sys, time, numpy as np
sys.path.insert(0, '.')
from oracle import pyoracle as po
from iris_lama_b200 import synth
name = sys.argv[1] if len(sys.argv) > 1 else "room"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 300
ds = synth.make_dataset(name, T)
runs = []
for sh in (0, 1, 7):
    s = po.Slam2D(po.SlamOptions.defaults(trans_thresh=0.05, rot_thresh=0.05), shuffle=sh)
    s.set_pose(*ds.truth[0])
    runs.append(s)
mism_total = 0
for t in range(T):
    for s in runs: s.update(ds.scans[t], ds.odom[t])
    if t % 10 == 0 or t == T-1:
        n, mn, mx = runs[0].dm_bounds()
        w, h = int(mx[0]-mn[0]), int(mx[1]-mn[1])
        e0 = runs[0].export_dm(mn[0], mn[1], w, h)
        for k, s in enumerate(runs[1:]):
            e = s.export_dm(mn[0], mn[1], w, h)
            d_sq = int(((e0['sqdist'] != e['sqdist']) | (e0['valid'] != e['valid'])).sum())
            d_off = int(((e0['ox'] != e['ox']) | (e0['oy'] != e['oy'])).sum())
            d_known = int((e0['known'] != e['known']).sum())
            mism_total += d_sq
            if d_sq or t % 100 == 0 or t == T-1:
                print(t, "run", k+1, "sq/valid mism", d_sq, "offset mism", d_off, "known mism", d_known, "valid cells", int(e0['valid'].sum()),
                      "pose diff", np.abs(runs[0].state()-s.state()).max())
print("TOTAL sq mismatches", mism_total, "counters", runs[0].counters()[1])
