"""ctypes binding of the CPU oracle (oracle/_build/liblama_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py -- never by the product package iris_lama_b200.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liblama_oracle.so")
_lib = None

c_dp = C.POINTER(C.c_double)
c_u32p = C.POINTER(C.c_uint32)


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
            os.path.getmtime(os.path.join(_HERE, f)) for f in ("lama_oracle.hpp", "oracle_capi.cpp", "Makefile")):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


class PFOptions(C.Structure):
    _fields_ = [("particles", C.c_uint32), ("srr", C.c_double), ("str", C.c_double), ("stt", C.c_double), ("srt", C.c_double),
                ("meas_sigma", C.c_double), ("meas_sigma_gain", C.c_double), ("trans_thresh", C.c_double), ("rot_thresh", C.c_double),
                ("l2_max", C.c_double), ("truncated_ray", C.c_double), ("truncated_range", C.c_double), ("resolution", C.c_double),
                ("patch_size", C.c_uint32), ("max_iter", C.c_uint32), ("threads", C.c_int32), ("seed", C.c_uint32)]

    @classmethod
    def defaults(cls, particles, **kw):
        # include/lama/pf_slam2d.h:132-185
        o = cls(particles=particles, srr=0.1, str=0.2, stt=0.1, srt=0.2, meas_sigma=0.05, meas_sigma_gain=3.0, trans_thresh=0.5,
                rot_thresh=0.5, l2_max=0.5, truncated_ray=0.0, truncated_range=0.0, resolution=0.05, patch_size=32, max_iter=100,
                threads=-1, seed=0)
        for k, v in kw.items():
            setattr(o, k, v)
        return o


class SlamOptions(C.Structure):
    _fields_ = [("trans_thresh", C.c_double), ("rot_thresh", C.c_double), ("l2_max", C.c_double), ("truncated_ray", C.c_double),
                ("truncated_range", C.c_double), ("resolution", C.c_double), ("patch_size", C.c_uint32), ("max_iter", C.c_uint32),
                ("strategy", C.c_int32), ("transient_map", C.c_int32)]

    @classmethod
    def defaults(cls, **kw):
        # include/lama/slam2d.h:91-125
        o = cls(trans_thresh=0.5, rot_thresh=0.5, l2_max=0.5, truncated_ray=0.0, truncated_range=0.0, resolution=0.05, patch_size=32,
                max_iter=100, strategy=0, transient_map=0)
        for k, v in kw.items():
            setattr(o, k, v)
        return o


class LocOptions(C.Structure):
    _fields_ = [("trans_thresh", C.c_double), ("rot_thresh", C.c_double), ("l2_max", C.c_double), ("resolution", C.c_double),
                ("patch_size", C.c_uint32), ("max_iter", C.c_uint32), ("strategy", C.c_int32), ("gloc_particles", C.c_uint32),
                ("gloc_iters", C.c_uint32), ("gloc_thresh", C.c_double), ("cov_blend", C.c_double)]

    @classmethod
    def defaults(cls, **kw):
        # src/loc2d.cpp:46-58
        o = cls(trans_thresh=0.5, rot_thresh=0.5, l2_max=1.0, resolution=0.05, patch_size=32, max_iter=100, strategy=0,
                gloc_particles=3000, gloc_iters=10, gloc_thresh=0.15, cov_blend=0.0)
        for k, v in kw.items():
            setattr(o, k, v)
        return o


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_rng_create.restype = C.c_void_p
        L.orc_rng_uniform.restype = C.c_double
        L.orc_rng_normal.restype = C.c_double
        L.orc_rng_raw.restype = C.c_uint32
        L.orc_se2_log_rot.restype = C.c_double
        L.orc_m2p.restype = C.c_uint64
        L.orc_m2c.restype = C.c_uint32
        for f in ("orc_ddm_create", "orc_ddm_clone", "orc_pf_create", "orc_pf_dm_handle", "orc_slam_create", "orc_slam_dm_handle", "orc_slamp_create",
                  "orc_loc_create", "orc_loc_dm_handle"):
            getattr(L, f).restype = C.c_void_p
        L.orc_ddm_update.restype = C.c_uint32
        L.orc_ddm_max_sqdist.restype = C.c_uint32
        L.orc_ddm_peak_queue.restype = C.c_uint64
        L.orc_pf_neff.restype = C.c_double
        _lib = L
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_dp)


def _u32(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a, a.ctypes.data_as(c_u32p)


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ---- Lie -------------------------------------------------------------------------------------------
def se2_from_xyr(x, y, r):
    out = np.zeros(4)
    lib().orc_se2_from_xyr(C.c_double(x), C.c_double(y), C.c_double(r), out.ctypes.data_as(c_dp))
    return out


def se2_exp(h):
    h, hp = _d(h)
    out = np.zeros(4)
    lib().orc_se2_exp(hp, out.ctypes.data_as(c_dp))
    return out


def se2_mul(a, b):
    a, ap = _d(a)
    b, bp = _d(b)
    out = np.zeros(4)
    lib().orc_se2_mul(ap, bp, out.ctypes.data_as(c_dp))
    return out


def se2_inv(a):
    a, ap = _d(a)
    out = np.zeros(4)
    lib().orc_se2_inv(ap, out.ctypes.data_as(c_dp))
    return out


def se2_rot(a):
    a, ap = _d(a)
    return lib().orc_se2_log_rot(ap)


class Rng:
    def __init__(self, seed):
        self.h = C.c_void_p(lib().orc_rng_create(C.c_uint32(seed)))

    def __del__(self):
        if self.h and _lib is not None:
            _lib.orc_rng_destroy(self.h)
            self.h = None

    def uniform(self):
        return lib().orc_rng_uniform(self.h)

    def normal(self, sigma):
        return lib().orc_rng_normal(self.h, C.c_double(sigma))

    def raw(self):
        return lib().orc_rng_raw(self.h)


# ---- addressing -----------------------------------------------------------------------------------
def w2m(p, res=0.05, patch=32):
    p, pp = _d(p)
    out = np.zeros(3, dtype=np.uint32)
    lib().orc_w2m(C.c_double(res), C.c_uint32(patch), pp, out.ctypes.data_as(c_u32p))
    return out


def w2m_nocast(p, res=0.05, patch=32):
    p, pp = _d(p)
    out = np.zeros(3)
    lib().orc_w2m_nocast(C.c_double(res), C.c_uint32(patch), pp, out.ctypes.data_as(c_dp))
    return out


def m2p(c, res=0.05, patch=32):
    c, cp = _u32(c)
    return lib().orc_m2p(C.c_double(res), C.c_uint32(patch), cp)


def m2c(c, res=0.05, patch=32):
    c, cp = _u32(c)
    return lib().orc_m2c(C.c_double(res), C.c_uint32(patch), cp)


def ray(a, b, cap=65536):
    a, ap = _u32(a)
    b, bp = _u32(b)
    out = np.zeros((cap, 3), dtype=np.uint32)
    n = lib().orc_ray(ap, bp, out.ctypes.data_as(c_u32p), C.c_int(cap))
    return out[:n].copy()


OFFSET = 1321122 * 32  # (UNIVERSAL_CONSTANT >> 1) * patch_length for 32-cell patches


def _export_dm(fn, args, x0, y0, w, h):
    out = dict(sqdist=np.zeros((h, w), np.uint16), valid=np.zeros((h, w), np.uint8), known=np.zeros((h, w), np.uint8),
               ox=np.zeros((h, w), np.int16), oy=np.zeros((h, w), np.int16), queued=np.zeros((h, w), np.uint8))
    fn(*args, C.c_uint32(x0), C.c_uint32(y0), C.c_int(w), C.c_int(h), _vp(out["sqdist"]), _vp(out["valid"]), _vp(out["known"]),
       _vp(out["ox"]), _vp(out["oy"]), _vp(out["queued"]))
    return out


def _export_occ(fn, args, x0, y0, w, h):
    out = dict(occupied=np.zeros((h, w), np.uint16), visited=np.zeros((h, w), np.uint16), known=np.zeros((h, w), np.uint8))
    fn(*args, C.c_uint32(x0), C.c_uint32(y0), C.c_int(w), C.c_int(h), _vp(out["occupied"]), _vp(out["visited"]), _vp(out["known"]))
    return out


def _bounds(fn, args):
    mn = np.zeros(2, np.uint32)
    mx = np.zeros(2, np.uint32)
    n = fn(*args, mn.ctypes.data_as(c_u32p), mx.ctypes.data_as(c_u32p))
    return n, mn, mx


class DDM:
    """Stand-alone DynamicDistanceMap (or a borrowed handle into a front end's map)."""

    def __init__(self, res=0.05, patch=32, l2_max=0.5, handle=None, owner=None):
        self.owner = owner
        self.owned = handle is None
        self.h = C.c_void_p(handle if handle is not None else lib().orc_ddm_create(C.c_double(res), C.c_uint32(patch), C.c_double(l2_max)))

    def __del__(self):
        if getattr(self, "owned", False) and self.h and _lib is not None:
            _lib.orc_ddm_destroy(self.h)
            self.h = None

    def clone(self):
        d = DDM.__new__(DDM)
        d.owner, d.owned = None, True
        d.h = C.c_void_p(lib().orc_ddm_clone(self.h))
        return d

    def set_shuffle(self, s):
        lib().orc_ddm_set_shuffle(self.h, C.c_uint32(s))

    @property
    def max_sqdist(self):
        return lib().orc_ddm_max_sqdist(self.h)

    @property
    def peak_queue(self):
        return lib().orc_ddm_peak_queue(self.h)

    def add(self, cells):
        c, cp = _u32(cells)
        lib().orc_ddm_add(self.h, cp, C.c_int(c.size // 2))

    def remove(self, cells):
        c, cp = _u32(cells)
        lib().orc_ddm_remove(self.h, cp, C.c_int(c.size // 2))

    def update(self):
        return lib().orc_ddm_update(self.h)

    def distance(self, pts, grad=True):
        p, pp = _d(pts)
        n = p.size // 3
        d = np.zeros(n)
        g = np.zeros((n, 3)) if grad else None
        lib().orc_ddm_distance(self.h, pp, C.c_int(n), d.ctypes.data_as(c_dp), g.ctypes.data_as(c_dp) if grad else None)
        return (d, g) if grad else d

    def distance_cells(self, cells):
        c, cp = _u32(cells)
        n = c.size // 2
        d = np.zeros(n)
        lib().orc_ddm_distance_cells(self.h, cp, C.c_int(n), d.ctypes.data_as(c_dp))
        return d

    def bounds(self):
        return _bounds(lib().orc_ddm_bounds, (self.h,))

    def export(self, x0, y0, w, h):
        return _export_dm(lib().orc_ddm_export, (self.h,), x0, y0, w, h)

    # ---- matching against this map ----
    def match_eval(self, pts, state, origin=(0, 0, 0), quat=(0, 0, 0, 1), jac=True):
        p, pp = _d(pts)
        n = p.size // 3
        o, op = _d(origin)
        q, qp = _d(quat)
        s, sp = _d(state)
        r = np.zeros(n)
        J = np.zeros((n, 3)) if jac else None
        lib().orc_match_eval(self.h, pp, C.c_int(n), op, qp, sp, r.ctypes.data_as(c_dp), J.ctypes.data_as(c_dp) if jac else None)
        return r, J

    def match_normal_eq(self, pts, state, robust=(1, 0.15), origin=(0, 0, 0), quat=(0, 0, 0, 1)):
        p, pp = _d(pts)
        o, op = _d(origin)
        q, qp = _d(quat)
        s, sp = _d(state)
        out = np.zeros(11)
        lib().orc_match_normal_eq(self.h, pp, C.c_int(p.size // 3), op, qp, sp, C.c_int(robust[0]), C.c_double(robust[1]), out.ctypes.data_as(c_dp))
        return out

    def match_error(self, pts, state, origin=(0, 0, 0), quat=(0, 0, 0, 1)):
        """MatchSurface2D::error (match_surface_2d.cpp:92-116)"""
        p, pp = _d(pts); o, op = _d(origin); q, qp = _d(quat); s, sp = _d(state)
        lib().orc_match_error.restype = C.c_double
        return lib().orc_match_error(self.h, pp, C.c_int(p.size // 3), op, qp, sp)

    def correlate_candidate_scan(self, pts, ref_xyr, cand_xyr, origin=(0, 0, 0), quat=(0, 0, 0, 1)):
        """GraphSlam2D::correlateCandidateScan (graph_slam2d.cpp:315-355) against this map -> (between xyr, rmse)"""
        p, pp = _d(pts); o, op = _d(origin); q, qp = _d(quat); r, rp = _d(ref_xyr); c, cp = _d(cand_xyr)
        out = np.zeros(3)
        lib().orc_correlate_candidate_scan.restype = C.c_double
        rmse = lib().orc_correlate_candidate_scan(self.h, pp, C.c_int(p.size // 3), op, qp, rp, cp, out.ctypes.data_as(c_dp))
        return out, rmse

    def coarse_correlate_candidate_scan(self, ref_pts, pts, ref_xyr, cand_xyr):
        """GraphSlam2D::coarseSearchAndCorrelateCandidateScan (graph_slam2d.cpp:357-392) -> (between xyr, rmse)"""
        a, ap = _d(ref_pts); p, pp = _d(pts); o, op = _d(_ID3); q, qp = _d(_IDQ); r, rp = _d(ref_xyr); c, cp = _d(cand_xyr)
        out = np.zeros(3)
        lib().orc_coarse_correlate_candidate_scan.restype = C.c_double
        rmse = lib().orc_coarse_correlate_candidate_scan(self.h, ap, C.c_int(a.size // 3), op, qp, pp, C.c_int(p.size // 3), op, qp, rp, cp, out.ctypes.data_as(c_dp))
        return out, rmse

    def match_solve(self, pts, state, strategy=0, robust=(1, 0.15), max_iter=100, want_cov=False, origin=(0, 0, 0), quat=(0, 0, 0, 1)):
        p, pp = _d(pts)
        o, op = _d(origin)
        q, qp = _d(quat)
        s = np.array(state, dtype=np.float64)
        cov = np.zeros((3, 3)) if want_cov else None
        stats = np.zeros(2, np.uint32)
        lib().orc_match_solve(self.h, pp, C.c_int(p.size // 3), op, qp, s.ctypes.data_as(c_dp), C.c_int(strategy), C.c_int(robust[0]),
                              C.c_double(robust[1]), C.c_uint32(max_iter), cov.ctypes.data_as(c_dp) if want_cov else None,
                              stats.ctypes.data_as(c_u32p))
        return s, cov, stats


_ID3 = np.zeros(3)
_IDQ = np.array([0.0, 0, 0, 1])


def loop_closure_candidates(key_xy, ignore_n, query, radius, max_candidates):
    """GraphSlam2D::findLoopClosureCandidates (graph_slam2d.cpp:283-313)"""
    k, kp = _d(key_xy); qq, qp = _d(query)
    ids = np.zeros(max(1, max_candidates), np.int32)
    n = lib().orc_loop_closure_candidates(kp, C.c_int(k.size // 2), C.c_int(ignore_n), qp, C.c_double(radius), C.c_int(max_candidates),
                                          ids.ctypes.data_as(C.POINTER(C.c_int32)))
    return ids[:n].copy()


class PFSlam2D:
    def __init__(self, opts: PFOptions, shuffle=0):
        self.opts = opts
        self.P = opts.particles
        self.h = C.c_void_p(lib().orc_pf_create(C.byref(opts)))
        if shuffle:
            lib().orc_pf_set_shuffle(self.h, C.c_uint32(shuffle))

    def __del__(self):
        if self.h and lib is not None and _lib is not None:
            _lib.orc_pf_destroy(self.h)
            self.h = None

    def set_prior(self, x, y, r):
        lib().orc_pf_set_prior(self.h, C.c_double(x), C.c_double(y), C.c_double(r))

    def update(self, pts, odom, origin=_ID3, quat=_IDQ):
        p, pp = _d(pts)
        o, op = _d(origin)
        q, qp = _d(quat)
        od, odp = _d(odom)
        return bool(lib().orc_pf_update(self.h, pp, C.c_int(p.size // 3), op, qp, odp))

    @property
    def neff(self):
        return lib().orc_pf_neff(self.h)

    def best(self):
        return lib().orc_pf_best(self.h)

    def particles(self):
        st = np.zeros((self.P, 4))
        w = np.zeros((self.P, 3))
        lib().orc_pf_get_particles(self.h, st.ctypes.data_as(c_dp), w.ctypes.data_as(c_dp))
        return st, w

    def last_resample(self):
        idx = np.zeros(self.P, np.int32)
        n = lib().orc_pf_last_resample(self.h, idx.ctypes.data_as(C.POINTER(C.c_int32)))
        return idx[:n].copy()

    def counters(self):
        last = np.zeros(6, np.uint64)
        tot = np.zeros(6, np.uint64)
        lib().orc_pf_counters(self.h, _vp(last), _vp(tot))
        keys = ("evals", "ray_cells", "dm_pops", "detached", "gn_iters", "resampled")
        return dict(zip(keys, last.tolist())), dict(zip(keys, tot.tolist()))

    def set_threads(self, n):
        lib().orc_pf_set_threads(self.h, C.c_int32(n))

    def memory_usage(self):
        """getMemoryUsage() and its (occmem, dmmem) overload: (total, occmem, dmmem)"""
        m = np.zeros(3, np.uint64)
        lib().orc_pf_memory_usage(self.h, m.ctypes.data_as(C.c_void_p))
        return int(m[0]), int(m[1]), int(m[2])

    def times(self):
        t = np.zeros(4)
        lib().orc_pf_times(self.h, t.ctypes.data_as(c_dp))
        return dict(zip(("solve", "normalize", "resample", "map"), t.tolist()))

    def trajectory(self, particle, cap=100000):
        out = np.zeros((cap, 3))
        n = lib().orc_pf_trajectory(self.h, C.c_int(particle), out.ctypes.data_as(c_dp), C.c_int(cap))
        return out[:n].copy()

    def dm_bounds(self, i):
        return _bounds(lib().orc_pf_dm_bounds, (self.h, C.c_int(i)))

    def occ_bounds(self, i):
        return _bounds(lib().orc_pf_occ_bounds, (self.h, C.c_int(i)))

    def export_dm(self, i, x0, y0, w, h):
        return _export_dm(lib().orc_pf_export_dm, (self.h, C.c_int(i)), x0, y0, w, h)

    def export_occ(self, i, x0, y0, w, h):
        return _export_occ(lib().orc_pf_export_occ, (self.h, C.c_int(i)), x0, y0, w, h)

    def dm(self, i):
        return DDM(handle=lib().orc_pf_dm_handle(self.h, C.c_int(i)), owner=self)


class Slam2D:
    def __init__(self, opts: SlamOptions, shuffle=0):
        self.opts = opts
        self.h = C.c_void_p(lib().orc_slam_create(C.byref(opts)))
        if shuffle:
            lib().orc_slam_set_shuffle(self.h, C.c_uint32(shuffle))

    def __del__(self):
        if self.h and _lib is not None:
            _lib.orc_slam_destroy(self.h)
            self.h = None

    def set_pose(self, x, y, r):
        lib().orc_slam_set_pose(self.h, C.c_double(x), C.c_double(y), C.c_double(r))

    def update(self, pts, odom, origin=_ID3, quat=_IDQ):
        p, pp = _d(pts)
        o, op = _d(origin)
        q, qp = _d(quat)
        od, odp = _d(odom)
        return bool(lib().orc_slam_update(self.h, pp, C.c_int(p.size // 3), op, qp, odp))

    def state(self):
        s = np.zeros(4)
        lib().orc_slam_get_state(self.h, s.ctypes.data_as(c_dp))
        return s

    def counters(self):
        last = np.zeros(6, np.uint64)
        tot = np.zeros(6, np.uint64)
        lib().orc_slam_counters(self.h, _vp(last), _vp(tot))
        keys = ("evals", "ray_cells", "dm_pops", "detached", "gn_iters", "resampled")
        return dict(zip(keys, last.tolist())), dict(zip(keys, tot.tolist()))

    def dm_bounds(self):
        return _bounds(lib().orc_slam_dm_bounds, (self.h,))

    def occ_bounds(self):
        return _bounds(lib().orc_slam_occ_bounds, (self.h,))

    def export_dm(self, x0, y0, w, h):
        return _export_dm(lib().orc_slam_export_dm, (self.h,), x0, y0, w, h)

    def export_occ(self, x0, y0, w, h):
        return _export_occ(lib().orc_slam_export_occ, (self.h,), x0, y0, w, h)

    def dm(self):
        return DDM(handle=lib().orc_slam_dm_handle(self.h), owner=self)


class Slam2DProb:
    """Slam2D over ProbabilisticOccupancyMap (log-odds cells)."""

    def __init__(self, opts: SlamOptions):
        self.opts = opts
        self.h = C.c_void_p(lib().orc_slamp_create(C.byref(opts)))

    def __del__(self):
        if self.h and _lib is not None:
            _lib.orc_slamp_destroy(self.h)
            self.h = None

    def set_pose(self, x, y, r):
        lib().orc_slamp_set_pose(self.h, C.c_double(x), C.c_double(y), C.c_double(r))

    def update(self, pts, odom, origin=_ID3, quat=_IDQ):
        p, pp = _d(pts)
        o, op = _d(origin)
        q, qp = _d(quat)
        od, odp = _d(odom)
        return bool(lib().orc_slamp_update(self.h, pp, C.c_int(p.size // 3), op, qp, odp))

    def state(self):
        s = np.zeros(4)
        lib().orc_slamp_get_state(self.h, s.ctypes.data_as(c_dp))
        return s

    def counters(self):
        last = np.zeros(6, np.uint64)
        lib().orc_slamp_counters(self.h, _vp(last))
        return dict(zip(("evals", "ray_cells", "dm_pops", "detached", "gn_iters", "resampled"), last.tolist()))

    def dm_bounds(self):
        return _bounds(lib().orc_slamp_dm_bounds, (self.h,))

    def occ_bounds(self):
        return _bounds(lib().orc_slamp_occ_bounds, (self.h,))

    def export_dm(self, x0, y0, w, h):
        return _export_dm(lib().orc_slamp_export_dm, (self.h,), x0, y0, w, h)

    def export_occ(self, x0, y0, w, h):
        out = dict(prob=np.zeros((h, w), np.float32), known=np.zeros((h, w), np.uint8))
        lib().orc_slamp_export_occ(self.h, C.c_uint32(x0), C.c_uint32(y0), C.c_int(w), C.c_int(h), _vp(out["prob"]), _vp(out["known"]))
        return out


def prob_sequence(kinds):
    k = np.ascontiguousarray(kinds, np.uint8)
    prob = np.zeros(len(k), np.float32)
    changed = np.zeros(len(k), np.uint8)
    lib().orc_prob_sequence(_vp(k), C.c_int(len(k)), _vp(prob), _vp(changed))
    return prob, changed


def prob_constants():
    out = np.zeros(5)
    lib().orc_prob_constants(out.ctypes.data_as(c_dp))
    return dict(zip(("miss", "hit", "clamp_min", "clamp_max", "occ_thresh"), out.tolist()))


class Loc2D:
    def __init__(self, opts: LocOptions):
        self.opts = opts
        self.h = C.c_void_p(lib().orc_loc_create(C.byref(opts)))

    def __del__(self):
        if self.h and _lib is not None:
            _lib.orc_loc_destroy(self.h)
            self.h = None

    def dm(self):
        return DDM(handle=lib().orc_loc_dm_handle(self.h), owner=self)

    def set_pose(self, x, y, r):
        lib().orc_loc_set_pose(self.h, C.c_double(x), C.c_double(y), C.c_double(r))

    def set_seed(self, seed):
        lib().orc_loc_set_seed(self.h, C.c_uint32(seed))

    def trigger_global_localization(self):
        lib().orc_loc_trigger_gloc(self.h)

    def gloc_active(self):
        return bool(lib().orc_loc_gloc_active(self.h))

    def occ_set(self, cells, state):
        c, cp = _u32(cells)
        lib().orc_loc_occ_set(self.h, cp, C.c_int(c.size // 2), C.c_int(state))

    def update(self, pts, odom, force=False, origin=_ID3, quat=_IDQ):
        p, pp = _d(pts)
        o, op = _d(origin)
        q, qp = _d(quat)
        od, odp = _d(odom)
        return bool(lib().orc_loc_update(self.h, pp, C.c_int(p.size // 3), op, qp, odp, C.c_int(int(force))))

    def get(self):
        s = np.zeros(4)
        cov = np.zeros((3, 3))
        rmse = C.c_double(0)
        stats = np.zeros(2, np.uint32)
        lib().orc_loc_get(self.h, s.ctypes.data_as(c_dp), cov.ctypes.data_as(c_dp), C.byref(rmse), stats.ctypes.data_as(c_u32p))
        return s, cov, rmse.value, stats


class LidarOdometry2D:
    """lama::LidarOdometry2D (src/lidar_odometry_2d.cpp:42-181)"""

    def __init__(self, resolution=0.05, max_iter=100):
        lib().orc_lo_create.restype = C.c_void_p
        self.h = C.c_void_p(lib().orc_lo_create(C.c_double(resolution), C.c_uint32(max_iter)))

    def __del__(self):
        if self.h and _lib is not None:
            _lib.orc_lo_destroy(self.h)
            self.h = None

    def update(self, pts, origin=_ID3, quat=_IDQ):
        p, pp = _d(pts)
        o, op = _d(origin)
        q, qp = _d(quat)
        return bool(lib().orc_lo_update(self.h, pp, C.c_int(p.size // 3), op, qp))

    def state(self):
        s = np.zeros(4)
        lib().orc_lo_get_state(self.h, s.ctypes.data_as(c_dp))
        return s

    def counters(self):
        c = np.zeros(6, np.uint64)
        lib().orc_lo_counters(self.h, _vp(c))
        return dict(zip(("evals", "ray_cells", "dm_pops", "gn_iters", "removed_patches", "map_updates"), c.tolist()))


# ---- SDM persistence and image export on map handles (oracle_capi.cpp) ------------------------------------------------
def _hv(h):
    h = getattr(h, "h", h)
    return h if isinstance(h, C.c_void_p) else C.c_void_p(h)


def map_handle(name, obj, *args):
    """name: pf_occ | pf_dm | slam_occ | slam_dm | slamp_occ | slamp_dm | loc_occ | loc_dm | lo_occ | lo_dm -> raw map pointer"""
    fn = getattr(lib(), "orc_%s_handle" % name)
    fn.restype = C.c_void_p
    return C.c_void_p(fn(_hv(obj), *[C.c_int(a) for a in args]))


def map_write(kind, handle, path):
    """kind: 'ddm' | 'freq' | 'prob' | 'simple' -- Map::write (map.cpp:490-529) of the map behind `handle`"""
    fn = getattr(lib(), "orc_%s_write" % kind)
    fn.restype = C.c_int
    return fn(_hv(handle), str(path).encode()) == 1


def map_read(kind, handle, path):
    """kind: 'ddm' | 'simple' -- Map::read (map.cpp:531-575)"""
    fn = getattr(lib(), "orc_%s_read" % kind)
    fn.restype = C.c_int
    return fn(_hv(handle), str(path).encode()) == 1


def map_image(kind, handle):
    """kind: 'ddm' | 'freq' | 'prob' -- the grey image of sdm::export_to_png (export.cpp:46-96) as a (height, width) array"""
    fn = getattr(lib(), "orc_%s_image" % kind)
    fn.restype = C.c_int
    dims = (C.c_int * 2)()
    fn(_hv(handle), None, C.c_size_t(0), dims)
    out = np.zeros((dims[1], dims[0]), np.uint8)
    if out.size:
        fn(_hv(handle), out.ctypes.data_as(C.c_void_p), C.c_size_t(out.size), dims)
    return out


def occ_query(kind, handle, cells):
    """kind 'freq' | 'prob': (getProbability, flags) of OccupancyMap for n cells; flags bit 0 isFree, 1 isOccupied, 2 isUnknown"""
    c = np.ascontiguousarray(cells, np.uint32).reshape(-1, 2)
    prob = np.zeros(len(c))
    flags = np.zeros(len(c), np.uint8)
    getattr(lib(), "orc_%s_query" % kind)(_hv(handle), c.ctypes.data_as(C.c_void_p), C.c_int(len(c)), prob.ctypes.data_as(C.c_void_p),
                                          flags.ctypes.data_as(C.c_void_p))
    return prob, flags
