"""Generates the golden fixtures under tests/golden/ from the CPU oracle (run from the repo root):

    python tests/golden/make_golden.py

The reference ships no golden vectors and cannot be built here (no Eigen), so these vectors come from the oracle
restatement; they pin (a) the oracle against regressions (-m "not gpu") and (b) the CUDA path without the oracle in
the loop (-m gpu).  Inputs are regenerated from iris_lama_b200.synth with fixed seeds."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from iris_lama_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
O = po.OFFSET

SLAM = dict(name="room", T=30, beams=360)
PF = dict(name="room", T=25, beams=180, P=12, seed=5, gain=0.02)
LOC = dict(T=4)


def room_cells(segments):
    cells = set()
    for x1, y1, x2, y2 in segments:
        n = int(max(abs(x2 - x1), abs(y2 - y1)) / 0.05) + 1
        for k in range(n + 1):
            cells.add((int((x1 + (x2 - x1) * k / n) * 20 + O + 0.5), int((y1 + (y2 - y1) * k / n) * 20 + O + 0.5)))
    return np.array(sorted(cells), np.uint32)


def golden_slam():
    ds = synth.make_dataset(SLAM["name"], SLAM["T"], n_beams=SLAM["beams"])
    s = po.Slam2D(po.SlamOptions.defaults(trans_thresh=0.05, rot_thresh=0.05))
    s.set_pose(*ds.truth[0])
    states, ctr = [], []
    for t in range(SLAM["T"]):
        s.update(ds.scans[t], ds.odom[t])
        states.append(s.state())
        c, _ = s.counters()
        ctr.append([c["evals"], c["ray_cells"], c["dm_pops"], c["gn_iters"]])
    n, mn, mx = s.dm_bounds()
    n2, mn2, mx2 = s.occ_bounds()
    lo = np.minimum(mn, mn2); hi = np.maximum(mx, mx2)
    w, h = int(hi[0] - lo[0]), int(hi[1] - lo[1])
    dm = s.export_dm(lo[0], lo[1], w, h)
    occ = s.export_occ(lo[0], lo[1], w, h)
    np.savez_compressed(os.path.join(HERE, "slam_room.npz"), states=np.array(states), counters=np.array(ctr, np.int64), origin=lo.astype(np.int64),
                        sqdist=dm["sqdist"], valid=dm["valid"], ox=dm["ox"].astype(np.int8), oy=dm["oy"].astype(np.int8), known=dm["known"],
                        occupied=occ["occupied"], visited=occ["visited"])


def golden_pf():
    ds = synth.make_dataset(PF["name"], PF["T"], n_beams=PF["beams"])
    o = po.PFSlam2D(po.PFOptions.defaults(PF["P"], trans_thresh=0.05, rot_thresh=0.05, seed=PF["seed"], meas_sigma_gain=PF["gain"]))
    o.set_prior(*ds.truth[0])
    resamples = -np.ones((PF["T"], PF["P"]), np.int32)
    neff, best = [], []
    for t in range(PF["T"]):
        o.update(ds.scans[t], ds.odom[t])
        idx = o.last_resample()
        if len(idx):
            resamples[t] = idx
        neff.append(o.neff)
        best.append(o.best())
    st, w = o.particles()
    n, mn, mx = o.occ_bounds(0)
    wd, hd = int(mx[0] - mn[0]), int(mx[1] - mn[1])
    occ = o.export_occ(3, mn[0], mn[1], wd, hd)
    dm = o.export_dm(3, mn[0], mn[1], wd, hd)
    np.savez_compressed(os.path.join(HERE, "pf_room.npz"), states=st, weights=w, resamples=resamples, neff=np.array(neff), best=np.array(best),
                        origin=mn.astype(np.int64), p3_visited=occ["visited"], p3_occupied=occ["occupied"], p3_sqdist=dm["sqdist"], p3_valid=dm["valid"],
                        trajectory0=o.trajectory(0))


def golden_loc():
    ds = synth.make_dataset("loc_room", LOC["T"])
    loc = po.Loc2D(po.LocOptions.defaults(trans_thresh=0.01, rot_thresh=0.01))
    d = loc.dm()
    d.add(room_cells(ds.segments))
    pops = d.update()
    t0 = ds.truth[0]
    loc.set_pose(t0[0] + 0.10, t0[1] - 0.07, t0[2] + 0.05)
    states, covs, rmses = [], [], []
    for t in range(LOC["T"]):
        loc.update(ds.scans[t], ds.odom[t], force=(t == 0))
        s, cov, rmse, _ = loc.get()
        states.append(s); covs.append(cov); rmses.append(rmse)
    np.savez_compressed(os.path.join(HERE, "loc_room.npz"), states=np.array(states), covs=np.array(covs), rmse=np.array(rmses), pops=np.array([pops]))


def golden_sdm():
    """a small distance map in the reference's on-disk format (Map::write, map.cpp:490-529): pins the byte layout"""
    d = po.DDM(0.05, 32, 0.5)
    cells = np.array([(O + 5, O + 7), (O + 6, O + 7), (O + 40, O - 3), (O - 20, O + 30)], np.uint32)
    d.add(cells)
    d.update()
    d.remove(cells[1:2])
    d.update()
    assert po.map_write("ddm", d, os.path.join(HERE, "ddm_small.sdm"))


if __name__ == "__main__":
    golden_slam()
    golden_pf()
    golden_loc()
    golden_sdm()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
