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
