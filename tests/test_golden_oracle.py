"""The oracle reproduces the committed golden fixtures (regression pin; generator: tests/golden/make_golden.py)."""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _gen():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_oracle_reproduces_slam_golden(po, synth):
    g = _gen()
    gold = np.load(os.path.join(HERE, "golden", "slam_room.npz"))
    ds = synth.make_dataset(g.SLAM["name"], g.SLAM["T"], n_beams=g.SLAM["beams"])
    s = po.Slam2D(po.SlamOptions.defaults(trans_thresh=0.05, rot_thresh=0.05))
    s.set_pose(*ds.truth[0])
    for t in range(g.SLAM["T"]):
        s.update(ds.scans[t], ds.odom[t])
        assert np.abs(s.state() - gold["states"][t]).max() < 1e-12
        c, _ = s.counters()
        assert [c["evals"], c["ray_cells"], c["dm_pops"], c["gn_iters"]] == gold["counters"][t].tolist()
    lo = gold["origin"]
    h, w = gold["sqdist"].shape
    dm = s.export_dm(int(lo[0]), int(lo[1]), w, h)
    occ = s.export_occ(int(lo[0]), int(lo[1]), w, h)
    assert (dm["sqdist"] == gold["sqdist"]).all() and (dm["valid"] == gold["valid"]).all() and (dm["known"] == gold["known"]).all()
    assert (dm["ox"] == gold["ox"]).all() and (dm["oy"] == gold["oy"]).all()
    assert (occ["occupied"] == gold["occupied"]).all() and (occ["visited"] == gold["visited"]).all()


def test_oracle_reproduces_pf_golden(po, synth):
    g = _gen()
    gold = np.load(os.path.join(HERE, "golden", "pf_room.npz"))
    c = g.PF
    ds = synth.make_dataset(c["name"], c["T"], n_beams=c["beams"])
    o = po.PFSlam2D(po.PFOptions.defaults(c["P"], trans_thresh=0.05, rot_thresh=0.05, seed=c["seed"], meas_sigma_gain=c["gain"]))
    o.set_prior(*ds.truth[0])
    n_res = 0
    for t in range(c["T"]):
        o.update(ds.scans[t], ds.odom[t])
        idx = o.last_resample()
        want = gold["resamples"][t]
        if want[0] >= 0:
            n_res += 1
            assert idx.tolist() == want.tolist()
        else:
            assert len(idx) == 0
        assert abs(o.neff - gold["neff"][t]) < 1e-9 and o.best() == gold["best"][t]
    assert n_res >= 1
    st, w = o.particles()
    assert np.abs(st - gold["states"]).max() < 1e-12 and np.abs(w - gold["weights"]).max() < 1e-9
