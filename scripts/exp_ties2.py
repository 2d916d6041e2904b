"""Random add/remove stress: literal vs shuffled ties vs brute-force truncated EDT."""
import sys, numpy as np
sys.path.insert(0, '.')
from oracle import pyoracle as po
from scipy import ndimage
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
O = po.OFFSET
W = 96
l2 = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
dms = [po.DDM(l2_max=l2) for _ in range(3)]
dms[1].set_shuffle(1); dms[2].set_shuffle(99)
occ = np.zeros((W, W), bool)
tot_mis = 0; tot_edt = 0; tot_cells = 0
for it in range(300):
    mode = rng.integers(0, 3)
    n = rng.integers(1, 40)
    if mode == 0:   # random points
        pts = rng.integers(16, W-16, size=(n, 2))
    elif mode == 1: # line segment
        x0, y0 = rng.integers(16, W-16, 2); dx, dy = rng.integers(-1, 2, 2)
        pts = np.array([(x0+k*dx, y0+k*dy) for k in range(n)]); pts = pts[(pts.min(1) >= 16) & (pts.max(1) < W-16)]
    else:
        pts = np.argwhere(occ)[:, ::-1]
        if len(pts): pts = pts[rng.choice(len(pts), size=min(len(pts), n), replace=False)]
    if len(pts) == 0: continue
    cells = (pts + O).astype(np.uint32)
    if mode == 2:
        for d in dms: d.remove(cells)
        occ[pts[:, 1], pts[:, 0]] = False
    else:
        for d in dms: d.add(cells)
        occ[pts[:, 1], pts[:, 0]] = True
    if rng.random() < 0.7:
        for d in dms: d.update()
        ex = [d.export(O, O, W, W) for d in dms]
        ms = dms[0].max_sqdist
        for k in (1, 2):
            tot_mis += int(((ex[0]['sqdist'] != ex[k]['sqdist']) | (ex[0]['valid'] != ex[k]['valid'])).sum())
        # brute force EDT
        if occ.any():
            edt2 = np.rint(ndimage.distance_transform_edt(~occ) ** 2).astype(np.int64)
        else:
            edt2 = np.full((W, W), 10**9)
        want_valid = edt2 < ms
        got_valid = ex[0]['valid'].astype(bool)
        bad = (want_valid != got_valid) | (want_valid & (edt2 != ex[0]['sqdist']))
        tot_edt += int(bad.sum()); tot_cells += W*W
print("tie mismatches", tot_mis, "vs exact EDT mismatches", tot_edt, "of", tot_cells)
