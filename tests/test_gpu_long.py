"""Long and sharded parity at the BASELINE.json size (256 particles x 1080 beams) and the corners of the map update the short
tests do not reach: hundreds of scans with several resamplings (copy-on-write detaches at full size), both forms of the ray cast,
G logical ranks on one device (pack / unpack / staging slots / local sources), a tilted sensor in the MAP UPDATE (3-axis walk,
map.cpp:198-227 with a moving z axis), and more candidate patches than the walk kernel has bitmaps for.
Match: src/pf_slam2d.cpp:178-312 (update), :254-302 (the two fan-outs), :537-574 (resample + COW copies)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
P, N = 256, 1080
POSE_TOL = 1e-9


def _cells_equal(g, o, particles):
    for p in particles:
        n, mn, mx = o.occ_bounds(p); w, h = int(mx[0] - mn[0]), int(mx[1] - mn[1])
        a, b = g.exportOccupancy(p, int(mn[0]), int(mn[1]), w, h), o.export_occ(p, mn[0], mn[1], w, h)
        assert (a["occupied"] == b["occupied"]).all() and (a["visited"] == b["visited"]).all() and (a["known"] == b["known"]).all(), p
        n, mn, mx = o.dm_bounds(p); w, h = int(mx[0] - mn[0]), int(mx[1] - mn[1])
        d1, d2 = g.exportDistance(p, int(mn[0]), int(mn[1]), w, h), o.export_dm(p, mn[0], mn[1], w, h)
        for k in ("sqdist", "valid", "ox", "oy", "queued", "known"):
            assert (d1[k] == d2[k]).all(), (p, k)


def _threads():
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        quota = int(q) // int(per) if q != "max" else 0
    except Exception:
        quota = 0
    n = len(os.sched_getaffinity(0))
    return max(2, min(n, quota) if quota else n)


@pytest.mark.parametrize("pull_max,T", [("48", 300), ("100000", 60)])
def test_fullsize_long_run_with_resampling(gpu_api, po, synth, monkeypatch, pull_max, T):
    """256 x 1080 over hundreds of scans, measurement gain low enough to resample every few dozen scans: particle states, weights,
    resampling history (exact), work counters and every cell of three particles against the oracle.  First with the default
    dispatch (256 particles -> the per-beam walk), then with the pull form of the ray cast forced at full size."""
    monkeypatch.setenv("LAMA_PULL_MAX_PARTICLES", pull_max)
    ds = synth.make_dataset("loop", T, n_beams=N)
    opts = dict(trans_thresh=0.05, rot_thresh=0.05, seed=42, meas_sigma_gain=0.0008)
    g = gpu_api.PFSlam2D(gpu_api.PFSlam2D.Options(P, **opts))
    o = po.PFSlam2D(po.PFOptions.defaults(P, threads=_threads(), **opts))
    g.setPrior(*ds.truth[0]); o.set_prior(*ds.truth[0])
    n_res, detached = 0, 0
    for t in range(T):
        assert g.update(ds.scans[t], ds.odom[t]) == o.update(ds.scans[t], ds.odom[t])
        rg, ro = g.lastResample(), o.last_resample()
        assert rg.tolist() == ro.tolist(), t
        n_res += int(len(ro) > 0)
        cg, _ = g.counters(); co, _ = o.counters()
        assert (cg["evals"], cg["gn_iters"]) == (co["evals"], co["gn_iters"]), t
        if len(ro) == 0:   # on resampling scans the device counts the map work of the set BEFORE the resampling (test_gpu_parity.py)
            assert (cg["ray_cells"], cg["dm_pops"]) == (co["ray_cells"], co["dm_pops"]), t
        if t % 25 == 0 or t == T - 1:
            sg, wg = g.getParticles(); so, wo = o.particles()
            assert np.abs(sg - so).max() < POSE_TOL, t
            assert np.abs(wg - wo).max() < 1e-6 * max(1.0, np.abs(wo).max()), t
    assert n_res >= (3 if T >= 300 else 2)
    _, tot = g.counters()
    assert tot["detached"] > 0                      # copy-on-write detaches really happened at full size
    _cells_equal(g, o, (0, 101, 255))
    assert g.getBestParticleIdx() == o.best()
    tg, to = g.trajectory(g.getBestParticleIdx()), o.trajectory(o.best())
    assert tg.shape == to.shape and np.abs(tg - to).max() < POSE_TOL


@pytest.mark.parametrize("G", [2, 4])
def test_logical_ranks_on_one_device_equal_the_single_process_oracle(gpu_api, po, synth, G):
    """G handles with shard_rank 0..G-1 on cuda:0, the exchange done by host copies (distributed.LocalShards):
    lama_pf_shard_begin / finish / apply_local / map_update and lama_pf_particle_pack / unpack with the staging slots, including
    resampling scans whose ancestors live on another rank (their updated maps migrate)."""
    from iris_lama_b200.distributed import LocalShards
    Pn, T, beams = 32, 40, 360
    ds = synth.make_dataset("room", T, n_beams=beams)
    opts = dict(trans_thresh=0.05, rot_thresh=0.05, seed=7, meas_sigma_gain=0.002)
    hs = [gpu_api.PFSlam2D(gpu_api.PFSlam2D.Options(Pn, shard_rank=r, shard_count=G, **opts)) for r in range(G)]
    o = po.PFSlam2D(po.PFOptions.defaults(Pn, threads=8, **opts))
    for h in hs:
        h.setPrior(*ds.truth[0])
    o.set_prior(*ds.truth[0])
    sh = LocalShards(hs, Pn)
    n_res = 0
    for t in range(T):
        assert sh.update(ds.scans[t], ds.odom[t]) == o.update(ds.scans[t], ds.odom[t])
        ro = o.last_resample()
        assert sh.last_idx.tolist() == ro.tolist(), t
        n_res += int(len(ro) > 0)
        so, wo = o.particles()
        for h in hs:                                     # every rank tracks all particle states / weights
            sg, wg = h.getParticles()
            assert np.abs(sg - so).max() < POSE_TOL and np.abs(wg - wo).max() < 1e-6 * max(1.0, np.abs(wo).max()), t
    assert n_res >= 2 and sh.migrated_bytes > 0          # maps really moved between ranks
    per = Pn // G

    class _View:   # particle p of the filter lives on rank p // per (the map getters of a sharded handle take the global index)
        def exportOccupancy(self, p, *a): return hs[p // per].exportOccupancy(p, *a)
        def exportDistance(self, p, *a): return hs[p // per].exportDistance(p, *a)
    _cells_equal(_View(), o, range(Pn))


def test_sharded_handle_needs_an_explicit_seed(gpu_api):
    with pytest.raises(gpu_api.LamaError):
        gpu_api.PFSlam2D(gpu_api.PFSlam2D.Options(8, shard_rank=0, shard_count=2, seed=0))


@pytest.mark.parametrize("pull_max", ["48", "0"])
def test_map_update_with_a_tilted_sensor(gpu_api, po, synth, monkeypatch, pull_max):
    """sensor pitched by 0.05 rad and mounted 0.3 m up: hit.z != start.z, so Map::computeRay's z axis moves (the 3-axis walk of
    k_raycast; k_ray_setup refuses such scans and hands the particle over) -- cells, counters and poses against the oracle"""
    monkeypatch.setenv("LAMA_PULL_MAX_PARTICLES", pull_max)
    Pn, T = 6, 12
    ds = synth.make_dataset("room", T, n_beams=360)
    origin = (0.1, 0.0, 0.3)
    a = 0.05
    quat = (0.0, np.sin(a / 2), 0.0, np.cos(a / 2))      # rotation about the y axis
    opts = dict(trans_thresh=0.05, rot_thresh=0.05, seed=3)
    g = gpu_api.PFSlam2D(gpu_api.PFSlam2D.Options(Pn, **opts))
    o = po.PFSlam2D(po.PFOptions.defaults(Pn, threads=4, **opts))
    g.setPrior(*ds.truth[0]); o.set_prior(*ds.truth[0])
    for t in range(T):
        assert g.update(ds.scans[t], ds.odom[t], origin=origin, quat=quat) == o.update(ds.scans[t], ds.odom[t], origin=origin, quat=quat)
        cg, _ = g.counters(); co, _ = o.counters()
        assert (cg["evals"], cg["ray_cells"], cg["dm_pops"], cg["gn_iters"]) == (co["evals"], co["ray_cells"], co["dm_pops"], co["gn_iters"]), t
    sg, _ = g.getParticles(); so, _ = o.particles()
    assert np.abs(sg - so).max() < POSE_TOL
    _cells_equal(g, o, range(Pn))


def test_more_candidate_patches_than_bitmaps(gpu_api, po, synth, monkeypatch):
    """LAMA_RAY_CAND_CAP = 2: every patch with hit cells or obstacles but two overflows the walk kernel's candidate bitmaps and
    falls back to logging every touch (kCandOverflow); the result must not change.  (Few beams: the log has to hold every touch
    of those patches; with a full scan the same setting ends in the loud LAMA_ERR_OVERFLOW instead, checked last.)"""
    monkeypatch.setenv("LAMA_PULL_MAX_PARTICLES", "0")
    monkeypatch.setenv("LAMA_RAY_CAND_CAP", "2")
    Pn, T = 4, 10
    ds = synth.make_dataset("room", T, n_beams=24)
    opts = dict(trans_thresh=0.05, rot_thresh=0.05, seed=11)
    g = gpu_api.PFSlam2D(gpu_api.PFSlam2D.Options(Pn, **opts))
    o = po.PFSlam2D(po.PFOptions.defaults(Pn, threads=4, **opts))
    g.setPrior(*ds.truth[0]); o.set_prior(*ds.truth[0])
    for t in range(T):
        assert g.update(ds.scans[t], ds.odom[t]) == o.update(ds.scans[t], ds.odom[t])
        cg, _ = g.counters(); co, _ = o.counters()
        assert (cg["ray_cells"], cg["dm_pops"]) == (co["ray_cells"], co["dm_pops"]), t
    _cells_equal(g, o, range(Pn))
    dense = synth.make_dataset("room", 3, n_beams=720)
    g2 = gpu_api.PFSlam2D(gpu_api.PFSlam2D.Options(Pn, **opts))
    g2.setPrior(*dense.truth[0])
    with pytest.raises(gpu_api.LamaError) as e:      # the touch log cannot hold a dense scan's touches of the overflowing patches: loud, never silent
        for t in range(3):
            g2.update(dense.scans[t], dense.odom[t])
        g2.getPose(); g2.counters()
    assert e.value.code == -6
