"""Independent pins of the CPU oracle (the reference ships no tests or golden vectors -- SURVEY.md section 4):
closed forms, numpy / scipy restatements and hand-enumerated cases."""
import math

import numpy as np
import pytest
from scipy import ndimage

O = 1321122 * 32


# ---- grid addressing (map.h:125-189, map.cpp:55-58) ---------------------------------------------------------
def test_w2m_round_trip_and_offset(po):
    assert po.OFFSET == O == 42275904
    assert (po.w2m([0.0, 0.0, 0.0]) == [O, O, O]).all()
    assert (po.w2m([0.024, -0.026, 0.0]) == [O, O - 1, O]).all()      # +0.5 then truncate
    assert (po.w2m([0.025, 0.0, 0.0])[0] in (O, O + 1))
    m = po.w2m_nocast([1.0, -2.5, 0.0])
    assert m[0] == 20.0 + O and m[1] == -50.0 + O
    c = np.array([O + 37, O - 5, O], np.uint32)
    assert po.m2p(c) == ((O + 37) >> 5) * 2642244 + ((O - 5) >> 5)
    assert po.m2c(c) == ((O + 37) & 31) | (((O - 5) & 31) << 5)


# ---- Bresenham (map.cpp:198-227): both endpoints excluded ----------------------------------------------------
def test_ray_hand_enumerated(po):
    a = np.array([O, O, O], np.uint32)
    assert len(po.ray(a, a)) == 0
    assert len(po.ray(a, a + np.array([1, 0, 0], np.uint32))) == 0            # |delta| = 1 -> nothing between
    r = po.ray(a, a + np.array([4, 0, 0], np.uint32))
    assert (r[:, 0] == [O + 1, O + 2, O + 3]).all() and (r[:, 1] == O).all()
    r = po.ray(a, (a.astype(np.int64) + [-3, -3, 0]).astype(np.uint32))
    assert (r[:, 0] == [O - 1, O - 2]).all() and (r[:, 1] == [O - 1, O - 2]).all()
    r = po.ray(a, a + np.array([5, 2, 0], np.uint32))                          # 2*err >= n rule
    assert r[:, 0].tolist() == [O + 1, O + 2, O + 3, O + 4]
    assert r[:, 1].tolist() == [O, O + 1, O + 1, O + 2]


def test_ray_properties(po):
    rng = np.random.default_rng(0)
    for _ in range(200):
        d = rng.integers(-60, 61, size=2)
        a = np.array([O + 5, O - 9, O], np.int64)
        b = a + [d[0], d[1], 0]
        r = po.ray(a.astype(np.uint32), b.astype(np.uint32)).astype(np.int64)
        n = max(abs(d[0]), abs(d[1]))
        assert len(r) == max(n - 1, 0)
        if len(r):
            steps = np.diff(np.vstack([a[None, :], r]), axis=0)
            assert np.abs(steps).max() <= 1
            assert not ((r == a).all(1).any() or (r == b).all(1).any())


# ---- SE2 / SO2 (sophus) --------------------------------------------------------------------------------------
def test_se2_identities(po):
    rng = np.random.default_rng(1)
    for _ in range(50):
        x, y, th = rng.uniform(-5, 5), rng.uniform(-5, 5), rng.uniform(-3.1, 3.1)
        s = po.se2_from_xyr(x, y, th)
        assert abs(s[0] - math.cos(th)) < 1e-15 and abs(s[1] - math.sin(th)) < 1e-15
        assert abs(po.se2_rot(s) - th) < 1e-14
        ident = po.se2_mul(s, po.se2_inv(s))
        assert np.abs(ident - [1, 0, 0, 0]).max() < 1e-14
        h = rng.uniform(-0.3, 0.3, 3)
        e = po.se2_exp(h)
        # closed form of the SE2 exponential
        a = math.sin(h[2]) / h[2]
        b = (1 - math.cos(h[2])) / h[2]
        assert abs(e[2] - (a * h[0] - b * h[1])) < 1e-15 and abs(e[3] - (b * h[0] + a * h[1])) < 1e-15
    e = po.se2_exp([0.3, -0.2, 1e-12])                                         # small-angle series branch
    assert abs(e[2] - 0.3) < 1e-12 and abs(e[3] + 0.2) < 1e-12


# ---- RNG: libstdc++ mt19937 == numpy MT19937 raw stream --------------------------------------------------------
def test_mt19937_matches_numpy(po):
    r = po.Rng(42)
    bg = np.random.MT19937()
    # numpy seeds MT19937(seed) through SeedSequence; set the legacy init_genrand state explicitly
    st = np.random.RandomState(42).get_state()
    bg.state = {"bit_generator": "MT19937", "state": {"key": st[1], "pos": st[2]}}
    want = bg.random_raw(8)
    got = [r.raw() for _ in range(8)]
    assert got == [int(v) for v in want]


def test_uniform_and_normal_are_libstdcxx(po):
    # generate_canonical<double,53> with a 32-bit engine: (lo + hi * 2^32) / 2^64
    r1, r2 = po.Rng(7), po.Rng(7)
    lo, hi = r2.raw(), r2.raw()
    assert r1.uniform() == (lo + hi * 4294967296.0) / 18446744073709551616.0
    # Marsaglia polar method as in libstdc++'s normal_distribution (fresh object per call: second variate dropped)
    r3, r4 = po.Rng(9), po.Rng(9)

    def canon():
        a, b = r4.raw(), r4.raw()
        return (a + b * 4294967296.0) / 18446744073709551616.0
    while True:
        x = 2.0 * canon() - 1.0
        y = 2.0 * canon() - 1.0
        r2_ = x * x + y * y
        if not (r2_ > 1.0 or r2_ == 0.0):
            break
    mult = math.sqrt(-2 * math.log(r2_) / r2_)
    assert r3.normal(0.37) == y * mult * 0.37 + 0.0


# ---- distance map ----------------------------------------------------------------------------------------------
def test_ddm_single_obstacle_is_exact_edt(po):
    for l2 in (0.5, 1.0):
        d = po.DDM(l2_max=l2)
        d.add(np.array([[O + 40, O + 40]], np.uint32))
        d.update()
        W = 80
        e = d.export(O, O, W, W)
        yy, xx = np.mgrid[0:W, 0:W]
        sq = (xx - 40) ** 2 + (yy - 40) ** 2
        want_valid = sq < d.max_sqdist
        assert (e["valid"].astype(bool) == want_valid).all()
        assert (e["sqdist"][want_valid] == sq[want_valid]).all()
        assert (e["ox"][want_valid] == (40 - xx)[want_valid]).all() and (e["oy"][want_valid] == (40 - yy)[want_valid]).all()
        assert e["queued"].sum() == 0


def test_ddm_walls_match_brute_force_edt(po):
    # straight walls + a box: on these the 4-neighbour brushfire is exact
    W = 96
    occ = np.zeros((W, W), bool)
    occ[20, 10:80] = True
    occ[20:70, 10] = True
    occ[50:56, 40:46] = True
    cells = np.argwhere(occ)[:, ::-1] + O
    d = po.DDM(l2_max=0.5)
    d.add(cells.astype(np.uint32))
    n = d.update()
    assert n > 0
    e = d.export(O, O, W, W)
    edt2 = np.rint(ndimage.distance_transform_edt(~occ) ** 2).astype(np.int64)
    want_valid = edt2 < d.max_sqdist
    assert (e["valid"].astype(bool) == want_valid).all()
    assert (e["sqdist"][want_valid] == edt2[want_valid]).all()
    # remove the box again: the distance field returns to the two-walls solution
    box = np.argwhere(occ[50:56, 40:46]) [:, ::-1] + [40 + O, 50 + O]
    d.remove(box.astype(np.uint32))
    d.update()
    occ[50:56, 40:46] = False
    e = d.export(O, O, W, W)
    edt2 = np.rint(ndimage.distance_transform_edt(~occ) ** 2).astype(np.int64)
    want_valid = edt2 < d.max_sqdist
    assert (e["valid"].astype(bool) == want_valid).all()
    assert (e["sqdist"][want_valid] == edt2[want_valid]).all()


def test_distance_gradient_finite_difference(po):
    d = po.DDM(l2_max=1.0)
    d.add(np.array([[O + 30 + k, O + 30] for k in range(30)], np.uint32))
    d.update()
    rng = np.random.default_rng(2)
    pts = np.zeros((100, 3))
    pts[:, 0] = rng.uniform(1.6, 2.9, 100)
    pts[:, 1] = rng.uniform(1.6, 2.4, 100)
    # keep away from cell borders where the bilinear gradient jumps
    m = pts[:, :2] * 20
    keep = (np.abs(m - np.rint(m)) > 0.05).all(1)
    pts = pts[keep]
    dist, grad = d.distance(pts)
    # h stays inside the cell (margin 0.05 cell = 2.5 mm); the absolute map coordinate (~4.2e7) only carries
    # ~7.5e-9 cells of precision, so a tiny h would be dominated by its quantisation
    h = 1e-3
    for k in range(2):
        p2 = pts.copy(); p2[:, k] += h
        p1 = pts.copy(); p1[:, k] -= h
        fd = (d.distance(p2, grad=False) - d.distance(p1, grad=False)) / (2 * h)
        assert np.abs(fd - grad[:, k]).max() < 1e-5
    # far from everything: max distance and zero gradient
    far, gfar = d.distance(np.array([[40.0, 40.0, 0.0]]))
    assert far[0] == math.sqrt(d.max_sqdist) * 0.05 and np.abs(gfar).max() == 0.0


# ---- matching: Gauss-Newton recovers an injected offset in an analytic room -------------------------------------------
def _room_dm(po, synth, l2=1.0):
    ds = synth.make_dataset("loc_room", 2)
    cells = set()
    for x1, y1, x2, y2 in ds.segments:
        n = int(max(abs(x2 - x1), abs(y2 - y1)) / 0.05) + 1
        for k in range(n + 1):
            cells.add((int((x1 + (x2 - x1) * k / n) * 20 + O + 0.5), int((y1 + (y2 - y1) * k / n) * 20 + O + 0.5)))
    d = po.DDM(l2_max=l2)
    d.add(np.array(sorted(cells), np.uint32))
    d.update()
    return ds, d


def test_gn_recovers_offset(po, synth):
    ds, d = _room_dm(po, synth)
    t = ds.truth[0]
    start = po.se2_from_xyr(t[0] + 0.10, t[1] - 0.07, t[2] + 0.05)
    for strategy in (0, 1):
        s, cov, stats = d.match_solve(ds.scans[0], start, strategy=strategy, want_cov=True)
        assert math.hypot(s[2] - t[0], s[3] - t[1]) < 0.01 and abs(po.se2_rot(s) - t[2]) < 0.005
        assert 1 <= stats[0] <= 30
        assert (np.linalg.eigvalsh(cov) > 0).all() and np.abs(cov - cov.T).max() < 1e-12
    # normal equations agree with a numpy restatement from residuals / Jacobian
    r, J = d.match_eval(ds.scans[0], start)
    w = np.sqrt(1.0 / (1.0 + r * r / 0.15 ** 2))
    rw, Jw = r * w, J * w[:, None]
    ne = d.match_normal_eq(ds.scans[0], start)
    A = Jw.T @ Jw
    assert np.allclose([A[0, 0], A[0, 1], A[0, 2], A[1, 1], A[1, 2], A[2, 2]], ne[:6], rtol=1e-12)
    assert np.allclose(Jw.T @ rw, ne[6:9], rtol=1e-12) and np.isclose(rw @ rw, ne[9], rtol=1e-12)
    # Jacobian vs finite differences of the left-perturbed residuals
    h = 1e-4
    for k in range(3):
        dv = np.zeros(3); dv[k] = h
        rp, _ = d.match_eval(ds.scans[0], po.se2_mul(po.se2_exp(dv), start), jac=False)
        rm, _ = d.match_eval(ds.scans[0], po.se2_mul(po.se2_exp(-dv), start), jac=False)
        fd = (rp - rm) / (2 * h)
        ok = np.abs(fd - J[:, k]) < 1e-3 * max(1.0, np.abs(J[:, k]).max())
        assert ok.mean() > 0.9           # the few outliers are endpoints that change cell within +-h


# ---- filter: normalise + systematic resampling vs a numpy restatement -------------------------------------------------
def test_pf_resampling_matches_numpy_restatement(po, synth):
    P, T = 12, 25
    ds = synth.make_dataset("room", T, n_beams=180)
    o = po.PFSlam2D(po.PFOptions.defaults(P, trans_thresh=0.05, rot_thresh=0.05, seed=5, meas_sigma_gain=0.02))
    o.set_prior(*ds.truth[0])
    seen = 0
    for t in range(T):
        _, w_before = o.particles()
        o.update(ds.scans[t], ds.odom[t])
        idx = o.last_resample()
        st, w = o.particles()
        if len(idx):
            seen += 1
            assert (np.diff(idx) >= 0).all() and idx.min() >= 0 and idx.max() < P
            assert (w[:, 0] == 0).all()                      # weight reset, weight_sum kept
    assert seen >= 1
    # numpy restatement of normalize() on the final weights
    st, w = o.particles()
    # deterministic: same seed, same data -> identical run
    o2 = po.PFSlam2D(po.PFOptions.defaults(P, trans_thresh=0.05, rot_thresh=0.05, seed=5, meas_sigma_gain=0.02))
    o2.set_prior(*ds.truth[0])
    for t in range(T):
        o2.update(ds.scans[t], ds.odom[t])
    st2, w2 = o2.particles()
    assert (st == st2).all() and (w == w2).all()


def test_systematic_resampler_restatement(po):
    # the exact arithmetic of pf_slam2d.cpp:537-553 in numpy on synthetic normalized weights
    rng = np.random.default_rng(3)
    for P in (8, 30, 256):
        w = rng.random(P); w /= w.sum()
        u = rng.random()
        interval = 1.0 / P
        target = interval * u
        cw, idx = 0.0, []
        for i in range(P):
            cw += w[i]
            while cw > target:
                idx.append(i); target += interval
        assert abs(len(idx) - P) <= 1
        counts = np.bincount(idx, minlength=P)
        assert (np.abs(counts - w * P) < 1.0 + 1e-9).all()


def test_thread_pool_equals_serial(po, synth):
    P, T = 8, 8
    ds = synth.make_dataset("room", T, n_beams=180)
    runs = []
    for threads in (-1, 4):
        o = po.PFSlam2D(po.PFOptions.defaults(P, trans_thresh=0.05, rot_thresh=0.05, seed=11, threads=threads))
        o.set_prior(*ds.truth[0])
        for t in range(T):
            o.update(ds.scans[t], ds.odom[t])
        runs.append(o.particles())
    assert (runs[0][0] == runs[1][0]).all() and (runs[0][1] == runs[1][1]).all()


def test_slam2d_tracks_truth(po, synth):
    T = 40
    ds = synth.make_dataset("room", T, n_beams=360)
    s = po.Slam2D(po.SlamOptions.defaults(trans_thresh=0.05, rot_thresh=0.05))
    s.set_pose(*ds.truth[0])
    for t in range(T):
        s.update(ds.scans[t], ds.odom[t])
    st = s.state()
    assert math.hypot(st[2] - ds.truth[-1, 0], st[3] - ds.truth[-1, 1]) < 0.05
    # gating: a scan without motion does not update
    assert s.update(ds.scans[-1], ds.odom[-1]) is False


# ---- independent pins of the "next" rows restated in round 1 ----------------------------------------
O_ = 1321122 * 32


def _wall_cells(segments):
    cells = set()
    for x1, y1, x2, y2 in segments:
        n = int(max(abs(x2 - x1), abs(y2 - y1)) / 0.05) + 1
        for k in range(n + 1):
            cells.add((int((x1 + (x2 - x1) * k / n) * 20 + O_ + 0.5), int((y1 + (y2 - y1) * k / n) * 20 + O_ + 0.5)))
    return np.array(sorted(cells), np.uint32)


def test_global_localization_finds_a_pose_that_explains_the_scan(po, synth):
    """Loc2D::globalLocalization (loc2d.cpp:249-286): from a wrong prior, the localiser must end at a pose whose scan lies on the
    mapped walls (the square room is symmetric, so the pose itself may be any of its four images)"""
    ds = synth.make_dataset("loc_room", 4)
    loc = po.Loc2D(po.LocOptions.defaults(trans_thresh=0.01, rot_thresh=0.01, gloc_particles=2000))
    cells = _wall_cells(ds.segments)
    d = loc.dm(); d.add(cells); d.update()
    xs = np.arange(int(-9.9 * 20), int(9.9 * 20), 2)
    loc.occ_set(np.array([(x + O_, y + O_) for x in xs for y in xs], np.uint32), -1)
    loc.set_seed(123)
    loc.set_pose(4.0, -3.0, 2.0)                       # far from the truth
    loc.trigger_global_localization()
    for t in range(4):
        loc.update(ds.scans[t], ds.odom[t], force=(t == 0))
    state, cov, rmse, _ = loc.get()
    assert rmse < 0.05 and not loc.gloc_active()
    c, s, tx, ty = state
    pts = ds.scans[3]
    wx, wy = c * pts[:, 0] - s * pts[:, 1] + tx, s * pts[:, 0] + c * pts[:, 1] + ty
    dist = np.minimum(np.minimum(np.abs(wx - 10), np.abs(wx + 10)), np.minimum(np.abs(wy - 10), np.abs(wy + 10)))   # 20 m square room
    assert np.median(dist) < 0.05


def test_sampling_covariance_is_a_covariance(po, synth):
    """Loc2D::addSamplingCovariance (loc2d.cpp:199-247) with cov_blend = 1: the xy block becomes the likelihood-weighted sample
    covariance of the 161 offset poses: symmetric, positive semi-definite, no wider than the sampled square (1 m)"""
    ds = synth.make_dataset("loc_room", 2)
    loc = po.Loc2D(po.LocOptions.defaults(trans_thresh=0.01, rot_thresh=0.01, cov_blend=1.0))
    d = loc.dm(); d.add(_wall_cells(ds.segments)); d.update()
    loc.set_pose(*ds.truth[0])
    loc.update(ds.scans[0], ds.odom[0], force=True)
    cov = loc.get()[1].reshape(3, 3)
    xy = cov[:2, :2]
    assert abs(xy[0, 1] - xy[1, 0]) < 1e-15 and (np.linalg.eigvalsh(xy) > -1e-12).all()
    assert 0 < xy[0, 0] < 1.0 and 0 < xy[1, 1] < 1.0
    # (an offset along a wall leaves the beams on that wall where they are, so the sum of per-beam likelihoods falls off slowly:
    #  the blended covariance is broad by construction, ~0.4 m here)


def test_lidar_odometry_tracks_and_keeps_a_local_map(po, synth):
    """LidarOdometry2D (lidar_odometry_2d.cpp:59-181): follows the motion without odometry; the transient map keeps only patches
    that meet the (pose-centred, 2 * maxDistance grown) AABB of the last mapped scan -- recomputed here independently"""
    ds = synth.make_dataset("room", 40, n_beams=1080)
    lo = po.LidarOdometry2D()
    for t in range(40):
        lo.update(ds.scans[t])
    c, s, tx, ty = lo.state()
    travelled = float(np.hypot(ds.truth[39][0] - ds.truth[0][0], ds.truth[39][1] - ds.truth[0][1]))
    assert abs(np.hypot(tx, ty) - travelled) < 0.05 * travelled and abs(np.arctan2(s, c)) < 0.02
    ds = synth.make_dataset("corridor", 60, n_beams=720)
    lo = po.LidarOdometry2D()
    near = lambda sc: np.ascontiguousarray(sc[np.hypot(sc[:, 0], sc[:, 1]) < 3.0])
    for t in range(60):
        lo.update(near(ds.scans[t]))
    assert lo.counters()["removed_patches"] > 0
    n, mn, mx = po._bounds(po.lib().orc_ddm_bounds, (po.map_handle("lo_dm", lo),))
    c, s, tx, ty = lo.state()
    # every surviving patch lies within (half scan extent <= 3 m) + 2 * maxDistance (2 m) + one patch (1.6 m) of the last mapped pose,
    # which is at most 0.1 m behind the current one
    lo_w, hi_w = (mn.astype(float) - O_) * 0.05, (mx.astype(float) - O_) * 0.05
    assert lo_w[0] > tx - 0.1 - 3.0 - 2.0 - 1.6 - 0.2 and hi_w[0] < tx + 0.1 + 3.0 + 2.0 + 1.6 + 0.2


def test_sdm_file_parsed_independently_equals_the_exported_planes(po, synth, tmp_path):
    """Map::write (map.cpp:490-529) read back with the numpy mirror (iris_lama_b200/sdm.py): cell for cell the exported map"""
    from iris_lama_b200 import sdm
    ds = synth.make_dataset("room", 5, n_beams=180)
    o = po.Slam2DProb(po.SlamOptions.defaults(trans_thresh=0.05, rot_thresh=0.05))
    o.set_pose(*ds.truth[0])
    for t in range(5):
        o.update(ds.scans[t], ds.odom[t])
    p = tmp_path / "occ.sdm"
    assert po.map_write("prob", po.map_handle("slamp_occ", o), p)
    f = sdm.read_sdm(p)
    assert f["header"]["cell_size"] == 4 and f["params"] == b""
    n, mn, mx = o.occ_bounds()
    e = o.export_occ(mn[0], mn[1], int(mx[0] - mn[0]), int(mx[1] - mn[1]))
    seen = 0
    for pid, (cells, mask) in f["patches"].items():
        x0, y0 = sdm.patch_origin(pid)
        bits = np.unpackbits(mask.view(np.uint8), bitorder="little").reshape(32, 32).astype(bool)     # bit i = cell (i & 31, i >> 5)
        sub = (slice(y0 - int(mn[1]), y0 - int(mn[1]) + 32), slice(x0 - int(mn[0]), x0 - int(mn[0]) + 32))
        assert (bits == e["known"][sub].astype(bool)).all()
        assert (cells.view("<f4").reshape(32, 32)[bits] == e["prob"][sub][bits]).all()
        seen += int(bits.sum())
    assert seen == int(e["known"].sum()) and len(f["patches"]) == n


def test_loop_closure_candidates_product_oracle_and_brute_force(po):
    """GraphSlam2D::findLoopClosureCandidates (graph_slam2d.cpp:283-313): the product's host function (no GPU needed), the oracle's restatement and a numpy
    brute force agree: first n - ignore key poses only, strictly inside the radius, nearest first, capped"""
    from iris_lama_b200 import api
    rng = np.random.default_rng(3)
    for trial in range(50):
        n = int(rng.integers(1, 200))
        keys = rng.uniform(-20, 20, size=(n, 2))
        q = rng.uniform(-20, 20, 2)
        radius = float(rng.uniform(0.5, 15))
        ignore = int(rng.integers(0, min(n, 25) + 1))
        cap = int(rng.integers(1, 8))
        d2 = ((keys[:n - ignore] - q) ** 2).sum(1)
        order = np.argsort(d2, kind="stable")
        want = [int(i) for i in order if d2[i] < radius * radius][:cap]
        assert api.loop_closure_candidates(keys, ignore, q, radius, cap).tolist() == want
        assert po.loop_closure_candidates(keys, ignore, q, radius, cap).tolist() == want
