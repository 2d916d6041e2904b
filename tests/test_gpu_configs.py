"""BASELINE.json configs 2 and 3 at their STATED sizes against the oracle (config 4 at 256 x 1080 x 5 000 scans takes minutes of CPU:
scripts/config4_parity.py, log under profiles/; its first 300 scans are tests/test_gpu_long.py; config 5 at 10 000 poses / 50 000 constraints is
tests/test_pgo.py::test_device_pgo_config5_size)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
POSE_TOL = 1e-9


def _cells_equal_slam(g, o):
    n, mn, mx = o.dm_bounds(); w, h = int(mx[0] - mn[0]), int(mx[1] - mn[1])
    a, b = g.exportDistance(int(mn[0]), int(mn[1]), w, h), o.export_dm(mn[0], mn[1], w, h)
    for k in ("sqdist", "valid", "ox", "oy", "queued", "known"):
        assert (a[k] == b[k]).all(), k
    n, mn, mx = o.occ_bounds(); w, h = int(mx[0] - mn[0]), int(mx[1] - mn[1])
    a, b = g.exportOccupancy(int(mn[0]), int(mn[1]), w, h), o.export_occ(mn[0], mn[1], w, h)
    for k in ("occupied", "visited", "known"):
        assert (a[k] == b[k]).all(), k
    return w * h


def test_config2_slam2d_corridor_720_beams_1000_scans(gpu_api, po, synth):
    """configs[1]: Slam2D online SLAM, 720-beam synthetic corridor, 1 000 scans, 0.05 m grid.  Every scan updates (thresholds below the 5 cm step);
    the 60 m corridor plus the 30 m lidar range needs the 128 x 128 patch window."""
    T = 1001                                                                   # scan 0 initialises the map, 1 000 scans follow
    ds = synth.make_dataset("corridor", T, n_beams=720)
    g = gpu_api.Slam2D(gpu_api.Slam2D.Options(trans_thresh=0.01, rot_thresh=0.01, dir_dim=128, max_beams=1024))   # 128 x 128 window: 204 m
    o = po.Slam2D(po.SlamOptions.defaults(trans_thresh=0.01, rot_thresh=0.01))
    g.setPose(*ds.truth[0]); o.set_pose(*ds.truth[0])
    ups = 0
    for t in range(T):
        a, b = g.update(ds.scans[t], ds.odom[t]), o.update(ds.scans[t], ds.odom[t])
        assert a == b
        ups += int(a)
        assert np.abs(g.state() - o.state()).max() < POSE_TOL, t
        if t % 50 == 0:
            cg, _ = g.counters(); co, _ = o.counters()
            assert (cg["evals"], cg["ray_cells"], cg["dm_pops"], cg["gn_iters"]) == (co["evals"], co["ray_cells"], co["dm_pops"], co["gn_iters"])
    assert ups >= 990
    _, tg = g.counters(); _, to = o.counters()
    assert all(tg[k] == to[k] for k in ("evals", "ray_cells", "dm_pops", "gn_iters"))
    assert _cells_equal_slam(g, o) > 50_000                                    # a 60 m x 2 m corridor with alcoves
    assert np.hypot(*(g.getPose()[:2] - ds.truth[-1, :2])) < 0.2               # and it is SLAM: the estimate stays on the true trajectory


def test_config3_pfslam2d_30_particles_1080_beams_2000_scans(gpu_api, po, synth):
    """configs[2]: PFSlam2D, 30 particles, 1080 beams, 0.05 m grid, 2 000-scan synthetic room (18 laps).  States, weights and resample indices
    on every scan; every cell of every particle at the end."""
    P, T = 30, 2001
    ds = synth.make_dataset("room", T, n_beams=1080)
    kw = dict(trans_thresh=0.05, rot_thresh=0.05, seed=42)
    g = gpu_api.PFSlam2D(gpu_api.PFSlam2D.Options(P, **kw))
    o = po.PFSlam2D(po.PFOptions.defaults(P, threads=8, **kw))
    g.setPrior(*ds.truth[0]); o.set_prior(*ds.truth[0])
    n_res = 0
    for t in range(T):
        assert g.update(ds.scans[t], ds.odom[t]) == o.update(ds.scans[t], ds.odom[t])
        rg, ro = g.lastResample(), o.last_resample()
        assert rg.tolist() == ro.tolist(), t
        n_res += int(len(ro) > 0)
        sg, wg = g.getParticles(); so, wo = o.particles()
        assert np.abs(sg - so).max() < POSE_TOL, t
        assert np.abs(wg - wo).max() < 1e-6 * max(1.0, np.abs(wo).max()), t
    assert g.getBestParticleIdx() == o.best()
    n, h = g.resampleDigest()
    assert n == n_res
    _, tg = g.counters(); _, to = o.counters()
    assert tg["evals"] == to["evals"] and tg["gn_iters"] == to["gn_iters"]
    cells = 0
    for p in range(P):
        _, mn, mx = o.occ_bounds(p); w, hh = int(mx[0] - mn[0]), int(mx[1] - mn[1])
        a, b = g.exportOccupancy(p, int(mn[0]), int(mn[1]), w, hh), o.export_occ(p, mn[0], mn[1], w, hh)
        assert (a["occupied"] == b["occupied"]).all() and (a["visited"] == b["visited"]).all() and (a["known"] == b["known"]).all(), p
        _, mn, mx = o.dm_bounds(p); w, hh = int(mx[0] - mn[0]), int(mx[1] - mn[1])
        a, b = g.exportDistance(p, int(mn[0]), int(mn[1]), w, hh), o.export_dm(p, mn[0], mn[1], w, hh)
        for k in ("sqdist", "valid", "ox", "oy", "queued", "known"):
            assert (a[k] == b[k]).all(), (p, k)
        cells += w * hh
    assert cells > 5_000_000
    assert np.hypot(*(g.getPose()[:2] - ds.truth[-1, :2])) < 0.2
