"""GPU parity tests proper: the CUDA path through the C-ABI against the oracle on the same seeded inputs and against
the committed golden fixtures.  Integer / cell data must be bit-exact; poses within 1e-9 (fp64 sums are reduced in a
different, fixed order on the device)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
O = 1321122 * 32
POSE_TOL = 1e-9          # BASELINE.json asks for 1e-4 on the trajectory; the device path is ~1e-12 in practice


def _room_cells(segments):
    cells = set()
    for x1, y1, x2, y2 in segments:
        n = int(max(abs(x2 - x1), abs(y2 - y1)) / 0.05) + 1
        for k in range(n + 1):
            cells.add((int((x1 + (x2 - x1) * k / n) * 20 + O + 0.5), int((y1 + (y2 - y1) * k / n) * 20 + O + 0.5)))
    return np.array(sorted(cells), np.uint32)


def _assert_dm_equal(a, b, fields=("sqdist", "valid", "ox", "oy", "queued", "known")):
    for k in fields:
        assert (a[k] == b[k]).all(), f"{k}: {int((a[k] != b[k]).sum())} cells differ"


# ---- DynamicDistanceMap grid interface ---------------------------------------------------------------------------------
@pytest.mark.parametrize("l2", [0.5, 1.0])
def test_brushfire_add_remove_stress_bit_exact(gpu_api, po, l2):
    g, o = gpu_api.DynamicDistanceMap(l2_max=l2), po.DDM(l2_max=l2)
    assert g.max_sqdist == o.max_sqdist
    rng = np.random.default_rng(11)
    W = 128
    occ = np.zeros((W, W), bool)
    for it in range(50):
        mode = rng.integers(0, 3)
        n = int(rng.integers(1, 60))
        if mode == 0:
            pts = rng.integers(24, W - 24, size=(n, 2))
        elif mode == 1:
            x0, y0 = rng.integers(24, W - 24, 2); dx, dy = rng.integers(-1, 2, 2)
            pts = np.array([(x0 + k * dx, y0 + k * dy) for k in range(n)]); pts = pts[(pts.min(1) >= 24) & (pts.max(1) < W - 24)]
        else:
            pts = np.argwhere(occ)[:, ::-1]
            if len(pts):
                pts = pts[rng.choice(len(pts), size=min(len(pts), n), replace=False)]
        if len(pts) == 0:
            continue
        cells = (pts + O).astype(np.uint32)
        if mode == 2:
            g.removeObstacle(cells); o.remove(cells); occ[pts[:, 1], pts[:, 0]] = False
        else:
            g.addObstacle(cells); o.add(cells); occ[pts[:, 1], pts[:, 0]] = True
        assert g.update() == o.update()             # DynamicDistanceMap::update() return value
        _assert_dm_equal(g.export(O, O, W, W), o.export(O, O, W, W))
    pts = np.zeros((2000, 3)); pts[:, :2] = rng.uniform(0.2, W * 0.05 - 0.2, size=(2000, 2))
    dg, gg = g.distance(pts); dc, gc = o.distance(pts)
    assert (dg == dc).all() and (gg == gc).all()   # same fp64 operations in the same order: bit-exact


def test_empty_and_degenerate_inputs(gpu_api, po):
    g, o = gpu_api.DynamicDistanceMap(), po.DDM()
    assert g.update() == o.update() == 0                                   # empty queues
    c = np.array([[O + 5, O + 5]], np.uint32)
    g.removeObstacle(c); o.remove(c)                                       # removing a non-obstacle is a no-op
    assert g.update() == o.update()
    g.addObstacle(np.vstack([c, c])); o.add(np.vstack([c, c]))             # duplicate add
    assert g.update() == o.update()
    g.addObstacle(c); g.removeObstacle(c); o.add(c); o.remove(c)           # add + remove before update
    assert g.update() == o.update()
    _assert_dm_equal(g.export(O - 32, O - 32, 96, 96), o.export(O - 32, O - 32, 96, 96))
    far = np.array([[100.0, 100.0, 0.0]])                                   # outside every patch: max distance, zero gradient
    d, gr = g.distance(far)
    assert d[0] == o.distance(far)[0][0] and np.abs(gr).max() == 0


def test_import_export_round_trip(gpu_api, po):
    o = po.DDM(l2_max=1.0)
    o.add(np.array([[O + 40 + k, O + 50] for k in range(40)] + [[O + 60, O + 20 + k] for k in range(30)], np.uint32))
    o.update()
    e = o.export(O, O, 128, 128)
    g = gpu_api.DynamicDistanceMap(l2_max=1.0)
    g.import_(O, O, e)
    _assert_dm_equal(g.export(O, O, 128, 128), e)
    # continuing on the imported map behaves like the oracle's own map
    rm = np.array([[O + 45, O + 50], [O + 60, O + 30]], np.uint32)
    g.removeObstacle(rm); o.remove(rm)
    assert g.update() == o.update()
    _assert_dm_equal(g.export(O, O, 128, 128), o.export(O, O, 128, 128))


# ---- MatchSurface2D / Solver -------------------------------------------------------------------------------------------
def test_normal_equations_and_solve(gpu_api, po, synth):
    ds = synth.make_dataset("loc_room", 2)
    cells = _room_cells(ds.segments)
    g, o = gpu_api.DynamicDistanceMap(l2_max=1.0), po.DDM(l2_max=1.0)
    g.addObstacle(cells); o.add(cells)
    assert g.update() == o.update()
    t = ds.truth[0]
    rng = np.random.default_rng(3)
    states = np.array([po.se2_from_xyr(t[0] + dx, t[1] + dy, t[2] + dth) for dx, dy, dth in
                       [(0.10, -0.07, 0.05), (0, 0, 0), (-0.2, 0.1, -0.08), (0.02, 0.01, 0.3)] + [tuple(rng.uniform(-0.15, 0.15, 3)) for _ in range(28)]])
    ne = g.matchNormalEquations(ds.scans[0], states)
    for i, s in enumerate(states):
        want = o.match_normal_eq(ds.scans[0], s)
        assert np.allclose(ne[i, :11], want, rtol=1e-11, atol=1e-13), i
    for strategy in (0, 1):
        sg, stats, sums = g.matchSolve(ds.scans[0], states, strategy=strategy)
        for i, s in enumerate(states):
            sc, _, st = o.match_solve(ds.scans[0], s, strategy=strategy)
            assert np.abs(sg[i] - sc).max() < POSE_TOL, (strategy, i)
            assert stats[i, 0] == st[0] and stats[i, 1] == st[1]          # same iteration / evaluation counts
    # robust kernels other than Cauchy and the max_iter edge cases
    for robust in ((0, 0.0), (2, 0.15), (3, 4.685), (4, 3.0)):
        ne = g.matchNormalEquations(ds.scans[0], states[:4], robust=robust)
        for i in range(4):
            assert np.allclose(ne[i, :11], o.match_normal_eq(ds.scans[0], states[i], robust=robust), rtol=1e-11, atol=1e-13)
    for max_iter in (0, 1, 2):
        sg, stats, _ = g.matchSolve(ds.scans[0], states[:4], max_iter=max_iter)
        for i in range(4):
            sc, _, st = o.match_solve(ds.scans[0], states[i], max_iter=max_iter)
            assert np.abs(sg[i] - sc).max() < POSE_TOL and stats[i, 0] == st[0]


def test_match_with_tilted_sensor_and_offset(gpu_api, po, synth):
    """general sensor pose: PointCloudXYZ::sensor_origin_ / sensor_orientation_ (types.h:117-118)"""
    ds = synth.make_dataset("loc_room", 1)
    cells = _room_cells(ds.segments)
    g, o = gpu_api.DynamicDistanceMap(l2_max=1.0), po.DDM(l2_max=1.0)
    g.addObstacle(cells); o.add(cells); g.update(); o.update()
    origin = (0.2, -0.1, 0.3)
    a = 0.1
    quat = (0.0, 0.0, np.sin(a / 2), np.cos(a / 2))
    t = ds.truth[0]
    s0 = po.se2_from_xyr(t[0] - 0.2, t[1] + 0.1, t[2] - a)
    ne = g.matchNormalEquations(ds.scans[0], [s0], origin=origin, quat=quat)
    assert np.allclose(ne[0, :11], o.match_normal_eq(ds.scans[0], s0, origin=origin, quat=quat), rtol=1e-11, atol=1e-13)


# ---- Loc2D (config 1) ---------------------------------------------------------------------------------------------------
def test_loc2d_matches_oracle_and_golden(gpu_api, po, synth):
    gold = np.load(os.path.join(HERE, "golden", "loc_room.npz"))
    ds = synth.make_dataset("loc_room", 4)
    cells = _room_cells(ds.segments)
    gl = gpu_api.Loc2D(gpu_api.Loc2D.Options(trans_thresh=0.01, rot_thresh=0.01))
    ol = po.Loc2D(po.LocOptions.defaults(trans_thresh=0.01, rot_thresh=0.01))
    gl.distance_map.addObstacle(cells)
    assert gl.distance_map.update() == int(gold["pops"][0])
    od = ol.dm(); od.add(cells); od.update()
    t0 = ds.truth[0]
    gl.setPose(t0[0] + 0.10, t0[1] - 0.07, t0[2] + 0.05); ol.set_pose(t0[0] + 0.10, t0[1] - 0.07, t0[2] + 0.05)
    for t in range(4):
        assert gl.update(ds.scans[t], ds.odom[t], force_update=(t == 0)) == ol.update(ds.scans[t], ds.odom[t], force=(t == 0))
        sc, cov, rmse, _ = ol.get()
        assert np.abs(gl.state() - sc).max() < POSE_TOL and np.abs(gl.state() - gold["states"][t]).max() < POSE_TOL
        assert abs(gl.getRMSE() - rmse) < 1e-12 and abs(gl.getRMSE() - gold["rmse"][t]) < 1e-12
        assert np.allclose(gl.getCovar(), cov, rtol=1e-9) and np.allclose(gl.getCovar(), gold["covs"][t], rtol=1e-9)
        assert np.hypot(sc[2] - ds.truth[t, 0], sc[3] - ds.truth[t, 1]) < 0.01
    assert gl.update(ds.scans[3], ds.odom[3]) is False                       # no motion -> gated


def test_loc2d_global_localization_and_sampling_covariance(gpu_api, po, synth):
    """Loc2D::globalLocalization (3000-candidate evaluation in one launch) + addSamplingCovariance"""
    ds = synth.make_dataset("loc_room", 6)
    cells = _room_cells(ds.segments)
    kw = dict(trans_thresh=0.01, rot_thresh=0.01, gloc_particles=3000, cov_blend=0.5)
    gl = gpu_api.Loc2D(gpu_api.Loc2D.Options(**kw))
    ol = po.Loc2D(po.LocOptions.defaults(**kw))
    gl.distance_map.addObstacle(cells); gl.distance_map.update()
    od = ol.dm(); od.add(cells); od.update()
    xs = np.arange(int(-9.9 * 20), int(9.9 * 20))
    free = np.array([(x + O, y + O) for x in xs[::2] for y in xs[::2]], np.uint32)
    gl.occupancySet(free, -1); gl.occupancySet(cells, 1)
    ol.occ_set(free, -1); ol.occ_set(cells, 1)
    gl.setSeed(77); ol.set_seed(77)
    gl.setPose(0, 0, 0); ol.set_pose(0, 0, 0)
    gl.triggerGlobalLocalization(); ol.trigger_global_localization()
    for t in range(6):
        assert gl.update(ds.scans[t], ds.odom[t], force_update=(t == 0)) == ol.update(ds.scans[t], ds.odom[t], force=(t == 0))
        sc, cov, rmse, _ = ol.get()
        assert np.abs(gl.state() - sc).max() < POSE_TOL, t              # same winning candidate, same refinement
        assert abs(gl.getRMSE() - rmse) < 1e-12
        assert np.allclose(gl.getCovar(), cov, rtol=1e-8, atol=1e-12)     # incl. the blended sampling covariance
        assert gl.globalLocalizationActive() == ol.gloc_active()
    assert gl.getRMSE() < 0.15 and not gl.globalLocalizationActive()


# ---- Slam2D (config 2 family) ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,beams,T", [("room", 360, 30), ("corridor", 720, 40), ("room", 1080, 12)])
def test_slam2d_matches_oracle(gpu_api, po, synth, name, beams, T):
    ds = synth.make_dataset(name, T, n_beams=beams)
    g = gpu_api.Slam2D(gpu_api.Slam2D.Options(trans_thresh=0.05, rot_thresh=0.05))
    o = po.Slam2D(po.SlamOptions.defaults(trans_thresh=0.05, rot_thresh=0.05))
    g.setPose(*ds.truth[0]); o.set_pose(*ds.truth[0])
    gold = np.load(os.path.join(HERE, "golden", "slam_room.npz")) if (name, beams, T) == ("room", 360, 30) else None
    for t in range(T):
        assert g.update(ds.scans[t], ds.odom[t]) == o.update(ds.scans[t], ds.odom[t])
        assert np.abs(g.state() - o.state()).max() < POSE_TOL
        cg, _ = g.counters(); co, _ = o.counters()
        assert (cg["evals"], cg["ray_cells"], cg["dm_pops"], cg["gn_iters"]) == (co["evals"], co["ray_cells"], co["dm_pops"], co["gn_iters"])
        assert g.getNumberOfProcessedCells() == co["dm_pops"]
        if gold is not None:
            assert np.abs(g.state() - gold["states"][t]).max() < POSE_TOL
            assert [cg["evals"], cg["ray_cells"], cg["dm_pops"], cg["gn_iters"]] == gold["counters"][t].tolist()
    n, mn, mx = o.dm_bounds(); w, h = int(mx[0] - mn[0]), int(mx[1] - mn[1])
    _assert_dm_equal(g.exportDistance(int(mn[0]), int(mn[1]), w, h), o.export_dm(mn[0], mn[1], w, h))
    ng, mng, mxg = g.mapBounds(1)
    assert (mng == mn).all() and (mxg == mx).all()                            # Map::bounds of the distance map
    n, mn, mx = o.occ_bounds(); w, h = int(mx[0] - mn[0]), int(mx[1] - mn[1])
    eg, eo = g.exportOccupancy(int(mn[0]), int(mn[1]), w, h), o.export_occ(mn[0], mn[1], w, h)
    for k in ("occupied", "visited", "known"):
        assert (eg[k] == eo[k]).all()
    ng, mng, mxg = g.mapBounds(0)
    assert ng == n and (mng == mn).all() and (mxg == mx).all()                # numOfPatches / bounds of the occupancy map
    if gold is not None:
        lo = gold["origin"]; hh, ww = gold["sqdist"].shape
        dg = g.exportDistance(int(lo[0]), int(lo[1]), ww, hh); og = g.exportOccupancy(int(lo[0]), int(lo[1]), ww, hh)
        assert (dg["sqdist"] == gold["sqdist"]).all() and (dg["valid"] == gold["valid"]).all() and (dg["known"] == gold["known"]).all()
        assert (og["occupied"] == gold["occupied"]).all() and (og["visited"] == gold["visited"]).all()


def test_slam2d_truncated_rays(gpu_api, po, synth):
    """Options::truncated_ray / truncated_range (slam2d.cpp:275-299)"""
    T = 10
    ds = synth.make_dataset("room", T, n_beams=360)
    for kw in (dict(truncated_ray=2.0), dict(truncated_range=6.0), dict(truncated_ray=1.5, truncated_range=8.0)):
        g = gpu_api.Slam2D(gpu_api.Slam2D.Options(trans_thresh=0.05, rot_thresh=0.05, **kw))
        o = po.Slam2D(po.SlamOptions.defaults(trans_thresh=0.05, rot_thresh=0.05, **kw))
        g.setPose(*ds.truth[0]); o.set_pose(*ds.truth[0])
        for t in range(T):
            assert g.update(ds.scans[t], ds.odom[t]) == o.update(ds.scans[t], ds.odom[t])
        assert np.abs(g.state() - o.state()).max() < POSE_TOL
        n, mn, mx = o.occ_bounds(); w, h = int(mx[0] - mn[0]), int(mx[1] - mn[1])
        eg, eo = g.exportOccupancy(int(mn[0]), int(mn[1]), w, h), o.export_occ(mn[0], mn[1], w, h)
        assert (eg["occupied"] == eo["occupied"]).all() and (eg["visited"] == eo["visited"]).all()


def test_slam2d_levenberg_marquardt(gpu_api, po, synth):
    T = 12
    ds = synth.make_dataset("room", T, n_beams=360)
    g = gpu_api.Slam2D(gpu_api.Slam2D.Options(trans_thresh=0.05, rot_thresh=0.05, strategy=1))
    o = po.Slam2D(po.SlamOptions.defaults(trans_thresh=0.05, rot_thresh=0.05, strategy=1))
    g.setPose(*ds.truth[0]); o.set_pose(*ds.truth[0])
    for t in range(T):
        g.update(ds.scans[t], ds.odom[t]); o.update(ds.scans[t], ds.odom[t])
        assert np.abs(g.state() - o.state()).max() < POSE_TOL


def test_slam2d_logodds_occupancy(gpu_api, po, synth):
    """row a18: ProbabilisticOccupancyMap cells (float log-odds, clamps) under the same update loop"""
    for name, beams, T in (("room", 360, 30), ("corridor", 720, 25)):
        ds = synth.make_dataset(name, T, n_beams=beams)
        g = gpu_api.Slam2D(gpu_api.Slam2D.Options(trans_thresh=0.05, rot_thresh=0.05, occupancy=1))
        o = po.Slam2DProb(po.SlamOptions.defaults(trans_thresh=0.05, rot_thresh=0.05))
        g.setPose(*ds.truth[0]); o.set_pose(*ds.truth[0])
        for t in range(T):
            assert g.update(ds.scans[t], ds.odom[t]) == o.update(ds.scans[t], ds.odom[t])
            assert np.abs(g.state() - o.state()).max() < POSE_TOL
            cg, _ = g.counters(); co = o.counters()
            assert (cg["evals"], cg["ray_cells"], cg["dm_pops"], cg["gn_iters"]) == (co["evals"], co["ray_cells"], co["dm_pops"], co["gn_iters"])
        n, mn, mx = o.occ_bounds(); w, h = int(mx[0] - mn[0]), int(mx[1] - mn[1])
        eg, eo = g.exportLogOdds(int(mn[0]), int(mn[1]), w, h), o.export_occ(mn[0], mn[1], w, h)
        assert (eg["prob"] == eo["prob"]).all()          # float cells bit-exact (same IEEE operations in the same order)
        assert (eg["known"] == eo["known"]).all()
        n, mn, mx = o.dm_bounds(); w, h = int(mx[0] - mn[0]), int(mx[1] - mn[1])
        _assert_dm_equal(g.exportDistance(int(mn[0]), int(mn[1]), w, h), o.export_dm(mn[0], mn[1], w, h))


# ---- PFSlam2D (config 3 family) -------------------------------------------------------------------------------------------
def _run_pf_pair(gpu_api, po, ds, P, T, **kw):
    g = gpu_api.PFSlam2D(gpu_api.PFSlam2D.Options(P, trans_thresh=0.05, rot_thresh=0.05, **kw))
    o = po.PFSlam2D(po.PFOptions.defaults(P, trans_thresh=0.05, rot_thresh=0.05, threads=8, **kw))
    g.setPrior(*ds.truth[0]); o.set_prior(*ds.truth[0])
    n_res = 0
    for t in range(T):
        assert g.update(ds.scans[t], ds.odom[t]) == o.update(ds.scans[t], ds.odom[t])
        rg, ro = g.lastResample(), o.last_resample()
        assert rg.tolist() == ro.tolist()                                     # resample indices bit-exact
        n_res += int(len(ro) > 0)
        sg, wg = g.getParticles(); so, wo = o.particles()
        assert np.abs(sg - so).max() < POSE_TOL
        assert np.abs(wg - wo).max() < 1e-6 * max(1.0, np.abs(wo).max())
        assert abs(g.getNeff() - o.neff) < 1e-6 and g.getBestParticleIdx() == o.best()
        cg, _ = g.counters(); co, _ = o.counters()
        assert (cg["evals"], cg["gn_iters"]) == (co["evals"], co["gn_iters"])
        if len(ro) == 0:
            # The device updates the maps BEFORE a resampling of the same scan and lets the offspring share the updated maps
            # (same maps as the reference's resample-then-update): on resampling scans its work counters refer to
            # the particle set before the resampling, the oracle's to the set after it.  The cells themselves are compared below.
            assert (cg["ray_cells"], cg["dm_pops"]) == (co["ray_cells"], co["dm_pops"])
    return g, o, n_res


def test_pfslam2d_with_forced_resampling_and_golden(gpu_api, po, synth):
    gold = np.load(os.path.join(HERE, "golden", "pf_room.npz"))
    P, T = 12, 25
    ds = synth.make_dataset("room", T, n_beams=180)
    g, o, n_res = _run_pf_pair(gpu_api, po, ds, P, T, seed=5, meas_sigma_gain=0.02)
    assert n_res >= 1                                                          # the COW / resample path really ran
    sg, wg = g.getParticles()
    assert np.abs(sg - gold["states"]).max() < POSE_TOL and np.abs(wg - gold["weights"]).max() < 1e-6
    lo = gold["origin"]; h, w = gold["p3_visited"].shape
    eg = g.exportOccupancy(3, int(lo[0]), int(lo[1]), w, h); dg = g.exportDistance(3, int(lo[0]), int(lo[1]), w, h)
    assert (eg["visited"] == gold["p3_visited"]).all() and (eg["occupied"] == gold["p3_occupied"]).all()
    assert (dg["sqdist"] == gold["p3_sqdist"]).all() and (dg["valid"] == gold["p3_valid"]).all()
    for p in range(P):
        n, mn, mx = o.dm_bounds(p); w, h = int(mx[0] - mn[0]), int(mx[1] - mn[1])
        _assert_dm_equal(g.exportDistance(p, int(mn[0]), int(mn[1]), w, h), o.export_dm(p, mn[0], mn[1], w, h))
        n, mn, mx = o.occ_bounds(p); w, h = int(mx[0] - mn[0]), int(mx[1] - mn[1])
        a, b = g.exportOccupancy(p, int(mn[0]), int(mn[1]), w, h), o.export_occ(p, mn[0], mn[1], w, h)
        assert (a["occupied"] == b["occupied"]).all() and (a["visited"] == b["visited"]).all() and (a["known"] == b["known"]).all()
    tg, to = g.trajectory(0), o.trajectory(0)
    assert tg.shape == to.shape and np.abs(tg - to).max() < POSE_TOL and np.abs(tg - gold["trajectory0"]).max() < POSE_TOL


def test_pfslam2d_config3_30_particles_1080_beams(gpu_api, po, synth):
    P, T = 30, 25
    ds = synth.make_dataset("room", T)
    assert ds.n_beams == 1080
    g, o, _ = _run_pf_pair(gpu_api, po, ds, P, T, seed=42)
    b = g.getBestParticleIdx()
    assert np.hypot(*(g.getPose()[:2] - ds.truth[T - 1, :2])) < 0.1
    n, mn, mx = o.dm_bounds(b); w, h = int(mx[0] - mn[0]), int(mx[1] - mn[1])
    _assert_dm_equal(g.exportDistance(b, int(mn[0]), int(mn[1]), w, h), o.export_dm(b, mn[0], mn[1], w, h))


def test_pfslam2d_motion_gate_and_rng_stream(gpu_api, po, synth):
    """default thresholds (0.5 m / 0.5 rad): most scans are gated but noise is still drawn on every call"""
    P, T = 8, 30
    ds = synth.make_dataset("room", T, n_beams=180)
    g = gpu_api.PFSlam2D(gpu_api.PFSlam2D.Options(P, seed=9))
    o = po.PFSlam2D(po.PFOptions.defaults(P, seed=9))
    g.setPrior(*ds.truth[0]); o.set_prior(*ds.truth[0])
    ups = 0
    for t in range(T):
        a, b = g.update(ds.scans[t], ds.odom[t]), o.update(ds.scans[t], ds.odom[t])
        assert a == b
        ups += int(a)
        assert np.abs(g.getParticles()[0] - o.particles()[0]).max() < POSE_TOL
    assert 2 <= ups < T


def test_pfslam2d_memory_usage_and_timestamps(gpu_api, po, synth):
    """getMemoryUsage (pf_slam2d.cpp:151-176 over Map::memory, map.cpp:115-125) and getTimestamps.  The occupancy maps have the reference's patches and,
    without resampling, its sharing state: byte counts agree (one truncation per map).  The reference's distance map also owns a patch wherever an
    occupancy cell was touched; the device keeps those cells in the occupancy patch and counts such a patch with THAT patch's use count, which the ray cast
    un-shares on any touch (the reference: on first touches): equal after the first scan (everything shared), an upper estimate afterwards."""
    P, T = 8, 10
    ds = synth.make_dataset("room", T, n_beams=180)
    g = gpu_api.PFSlam2D(gpu_api.PFSlam2D.Options(P, trans_thresh=0.05, rot_thresh=0.05, seed=5))
    o = po.PFSlam2D(po.PFOptions.defaults(P, trans_thresh=0.05, rot_thresh=0.05, seed=5))
    g.setPrior(*ds.truth[0]); o.set_prior(*ds.truth[0])
    assert g.getMemoryUsage() == (0, 0, 0) and g.getTimestamps() == []
    resampled, checked = False, 0
    for t in range(T):
        g.update(ds.scans[t], ds.odom[t], timestamp=100.0 + t); o.update(ds.scans[t], ds.odom[t])
        resampled |= len(o.last_resample()) > 0
        mg, mo = g.getMemoryUsage(), o.memory_usage()
        if resampled:                                                          # from here on the device store shares more than the reference
            assert 0 < mg[0] <= 1.5 * mo[0]
            continue
        checked += 1
        assert abs(mg[1] - mo[1]) <= 2 * P, (t, mg, mo)                       # occmem
        if t == 0:
            assert all(abs(a - b) <= 2 * P for a, b in zip(mg, mo)), (mg, mo)
        else:
            assert 0.8 * mo[2] <= mg[2] <= 1.3 * mo[2] and 0.8 * mo[0] <= mg[0] <= 1.3 * mo[0], (t, mg, mo)
    assert checked >= 2
    assert g.getTimestamps() == [100.0]
    g2, o2, n_res = _run_pf_pair(gpu_api, po, ds, P, T, seed=5, meas_sigma_gain=0.02)
    assert n_res >= 2
    assert 0 < g2.getMemoryUsage()[1] <= o2.memory_usage()[1] + 2 * P           # after a resampling the device store shares more


# ---- SURVEY 8(f) row 4 (front-end half): LidarOdometry2D and Slam2D's transient map ---------------------------------------
def _near(scan, r):
    """the beams shorter than r: a short-range sensor, so that the surface AABB moves with the robot"""
    return np.ascontiguousarray(scan[np.hypot(scan[:, 0], scan[:, 1]) < r])


def test_lidar_odometry_2d(gpu_api, po, synth):
    """LidarOdometry2D::update (lidar_odometry_2d.cpp:59-181): match, 1 m rays into the log-odds map, transient-map pruning"""
    ds = synth.make_dataset("corridor", 60, n_beams=720)
    g, o = gpu_api.LidarOdometry2D(), po.LidarOdometry2D()
    for t in range(60):
        s = _near(ds.scans[t], 3.0)
        assert g.update(s) == o.update(s)
        assert np.abs(g.state() - o.state()).max() < POSE_TOL, t
    c = o.counters()
    assert g.mapStats() == (c["map_updates"], c["removed_patches"]) and c["removed_patches"] > 0 and c["map_updates"] > 20
    n, mn, mx = po._bounds(po.lib().orc_ddm_bounds, (po.map_handle("lo_dm", o),))
    gb = g.mapBounds(1)
    assert (gb[1] == mn).all() and (gb[2] == mx).all()                    # the pruned map covers the same patches
    w, h = int(mx[0] - mn[0]), int(mx[1] - mn[1])
    _assert_dm_equal(g.exportDistance(mn[0], mn[1], w, h), po._export_dm(po.lib().orc_ddm_export, (po.map_handle("lo_dm", o),), mn[0], mn[1], w, h))
    assert (g.exportImage(0) == po.map_image("prob", po.map_handle("lo_occ", o))).all()
    assert g.getPose()[0] > 3.0                                            # it follows the forward motion (short-range beams in a corridor: weakly observable)


def test_slam2d_transient_map(gpu_api, po, synth):
    """Slam2D::Options::transient_map (slam2d.cpp:323-379)"""
    ds = synth.make_dataset("corridor", 90, n_beams=720)
    kw = dict(trans_thresh=0.05, rot_thresh=0.05, transient_map=1)
    g, o = gpu_api.Slam2D(gpu_api.Slam2D.Options(**kw)), po.Slam2D(po.SlamOptions.defaults(**kw))
    g.setPose(*ds.truth[0]); o.set_pose(*ds.truth[0])
    for t in range(90):
        s = _near(ds.scans[t], 1.5)
        assert g.update(s, ds.odom[t]) == o.update(s, ds.odom[t])
        assert np.abs(g.state() - o.state()).max() < POSE_TOL, t
    po.lib().orc_slam_removed_patches.restype = po.C.c_uint64
    removed = po.lib().orc_slam_removed_patches(o.h)
    assert removed > 0 and g.mapStats()[1] == removed
    n, mn, mx = o.dm_bounds()
    w, h = int(mx[0] - mn[0]), int(mx[1] - mn[1])
    _assert_dm_equal(g.exportDistance(mn[0], mn[1], w, h), o.export_dm(mn[0], mn[1], w, h))
    go, oo = g.exportOccupancy(mn[0], mn[1], w, h), o.export_occ(mn[0], mn[1], w, h)
    for k in ("occupied", "visited", "known"):
        assert (go[k] == oo[k]).all(), k


# ---- failures are loud (no silent truncation, no fallback) ------------------------------------------------------------------
def test_errors_are_reported_not_swallowed(gpu_api, synth):
    ds = synth.make_dataset("room", 4, n_beams=360)
    # a directory window of 8 x 8 patches (12.8 m) cannot hold a 20 m room: LAMA_ERR_WINDOW
    g = gpu_api.Slam2D(gpu_api.Slam2D.Options(trans_thresh=0.05, rot_thresh=0.05, dir_dim=8))
    g.setPose(*ds.truth[0])
    with pytest.raises(gpu_api.LamaError) as e:
        for t in range(3):
            g.update(ds.scans[t], ds.odom[t])
    assert e.value.code == -4 and "window" in str(e.value)
    # a pool of 24 patches cannot hold the first scan of the same room: LAMA_ERR_POOL
    g = gpu_api.Slam2D(gpu_api.Slam2D.Options(trans_thresh=0.05, rot_thresh=0.05, pool_slots=24))
    g.setPose(*ds.truth[0])
    with pytest.raises(gpu_api.LamaError) as e:
        for t in range(3):
            g.update(ds.scans[t], ds.odom[t])
    assert e.value.code == -5
    # the particle filter runs its map update asynchronously: the failure surfaces at the next call into the handle
    pf = gpu_api.PFSlam2D(gpu_api.PFSlam2D.Options(4, trans_thresh=0.05, rot_thresh=0.05, seed=3, dir_dim=8))
    pf.setPrior(*ds.truth[0])
    with pytest.raises(gpu_api.LamaError) as e:
        for t in range(4):
            pf.update(ds.scans[t], ds.odom[t])
        pf.counters()
    assert e.value.code == -4
    # more beams than the engine was created for
    g = gpu_api.Slam2D(gpu_api.Slam2D.Options(trans_thresh=0.05, rot_thresh=0.05, max_beams=360))
    g.setPose(*ds.truth[0])
    g.update(ds.scans[0], ds.odom[0])
    with pytest.raises(gpu_api.LamaError):
        g.update(np.zeros((400, 3)), ds.odom[1])


def test_occupancy_grid_queries(gpu_api, po, synth):
    """OccupancyMap::{getProbability,isFree,isOccupied,isUnknown} on device maps (occupancy_map.h:57-76), both map kinds"""
    ds = synth.make_dataset("room", 10, n_beams=360)
    rng = np.random.default_rng(2)
    for occupancy in (0, 1):
        kw = dict(trans_thresh=0.05, rot_thresh=0.05)
        g = gpu_api.Slam2D(gpu_api.Slam2D.Options(occupancy=occupancy, **kw))
        o = (po.Slam2DProb if occupancy else po.Slam2D)(po.SlamOptions.defaults(**kw))
        g.setPose(*ds.truth[0]); o.set_pose(*ds.truth[0])
        for t in range(10):
            g.update(ds.scans[t], ds.odom[t]); o.update(ds.scans[t], ds.odom[t])
        pts = np.c_[rng.uniform(-12, 12, (4000, 2)), np.zeros(4000)]          # inside, on and outside the mapped area
        cells = gpu_api.w2m(0.05, pts)
        assert (cells[:50] == np.array([po.w2m(p)[:2] for p in pts[:50]], np.uint32)).all()                 # Map::w2m
        dg, gg = g.distance(pts)                                               # getDistanceMap()->distance(point, &gradient)
        dr, gr = po.DDM(handle=po.map_handle("slamp_dm" if occupancy else "slam_dm", o).value, owner=o).distance(pts)
        assert (dg == dr).all() and (gg == gr).all()
        pg, fg = g.occupancyQuery(cells)
        pr, fr = po.occ_query("prob" if occupancy else "freq", po.map_handle("slamp_occ" if occupancy else "slam_occ", o), cells)
        assert (fg == fr).all() and (pg == pr).all()
        assert (fr & 1).any() and (fr & 2).any() and (fr & 4).any()            # free, occupied and unknown cells all occur


# ---- SURVEY 8(f) row 4 (second half): GraphSlam2D's loop-closure scan correlation -----------------------------------------------------
def test_loop_closure_scan_correlation(gpu_api, po, synth):
    """correlateCandidateScan / coarseSearchAndCorrelateCandidateScan (src/graph_slam2d.cpp:315-392) and MatchSurface2D::error
    (match_surface_2d.cpp:92-116) on a device distance map against the oracle: same `between` pose, same RMSE, for candidates that
    are close to, and far from, the reference key pose"""
    ds = synth.make_dataset("loc_room", 6)
    cells = _room_cells(ds.segments)
    g, o = gpu_api.DynamicDistanceMap(l2_max=0.5), po.DDM(l2_max=0.5)
    g.addObstacle(cells); o.add(cells); g.update(); o.update()
    rng = np.random.default_rng(8)
    for t in range(6):
        truth = ds.truth[t]
        st = po.se2_from_xyr(*(truth + rng.normal(0, [0.03, 0.03, 0.01])))
        assert abs(g.matchError(ds.scans[t], [st])[0] - o.match_error(ds.scans[t], st)) < 1e-12
        for dpose in ([0.05, -0.04, 0.02], [0.6, 0.3, -0.05], [-1.5, 0.8, 0.3]):
            cand = truth + np.array(dpose)                         # where the pose graph believes the candidate key pose is
            ref = ds.truth[(t + 2) % 6] + np.array([0.02, 0.01, -0.01])
            bg, eg = g.correlateCandidateScan(ds.scans[t], ref, cand)
            bo, eo = o.correlate_candidate_scan(ds.scans[t], ref, cand)
            assert np.abs(bg - bo).max() < POSE_TOL and abs(eg - eo) < 1e-9, (t, dpose)
            bg, eg = g.coarseCorrelateCandidateScan(ds.scans[(t + 2) % 6], ds.scans[t], ref, cand)
            bo, eo = o.coarse_correlate_candidate_scan(ds.scans[(t + 2) % 6], ds.scans[t], ref, cand)
            assert np.abs(bg - bo).max() < POSE_TOL and abs(eg - eo) < 1e-9, (t, dpose, "coarse")
    # a well-placed candidate ends on the walls: RMSE of the order of the range noise
    bg, eg = g.correlateCandidateScan(ds.scans[0], ds.truth[1], ds.truth[0] + np.array([0.08, -0.05, 0.03]))
    assert eg < 0.05
