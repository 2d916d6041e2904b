"""First-contact GPU parity script: prints a compact report per section, never stops at the first failure."""
import sys, time, traceback
import numpy as np
sys.path.insert(0, '.')
from oracle import pyoracle as po
from iris_lama_b200 import api, synth

O = api.OFFSET


def section(name):
    def deco(f):
        def run():
            t0 = time.time()
            try:
                f()
                print(f"[{name}] done in {time.time()-t0:.2f}s", flush=True)
            except Exception:
                print(f"[{name}] EXCEPTION\n{traceback.format_exc()}", flush=True)
        return run
    return deco


def cmp_dm(a, b, tag):
    out = []
    for k in ("sqdist", "valid", "ox", "oy", "queued", "known"):
        out.append(f"{k}:{int((a[k] != b[k]).sum())}")
    print(f"  {tag} mismatches {' '.join(out)} (valid cells {int(a['valid'].sum())})", flush=True)
    return int(((a["sqdist"] != b["sqdist"]) | (a["valid"] != b["valid"])).sum())


def cmp_occ(a, b, tag):
    out = [f"{k}:{int((a[k] != b[k]).sum())}" for k in ("occupied", "visited", "known")]
    print(f"  {tag} mismatches {' '.join(out)} (visited cells {int((a['visited'] > 0).sum())})", flush=True)


@section("ddm")
def t_ddm():
    rng = np.random.default_rng(0)
    W = 128
    for l2 in (0.5, 1.0):
        g = api.DynamicDistanceMap(l2_max=l2)
        o = po.DDM(l2_max=l2)
        occ = np.zeros((W, W), bool)
        for it in range(40):
            mode = rng.integers(0, 3)
            n = int(rng.integers(1, 60))
            if mode == 0:
                pts = rng.integers(24, W - 24, size=(n, 2))
            elif mode == 1:
                x0, y0 = rng.integers(24, W - 24, 2); dx, dy = rng.integers(-1, 2, 2)
                pts = np.array([(x0 + k * dx, y0 + k * dy) for k in range(n)]); pts = pts[(pts.min(1) >= 24) & (pts.max(1) < W - 24)]
            else:
                pts = np.argwhere(occ)[:, ::-1]
                if len(pts): pts = pts[rng.choice(len(pts), size=min(len(pts), n), replace=False)]
            if len(pts) == 0: continue
            cells = (pts + O).astype(np.uint32)
            if mode == 2:
                g.removeObstacle(cells); o.remove(cells); occ[pts[:, 1], pts[:, 0]] = False
            else:
                g.addObstacle(cells); o.add(cells); occ[pts[:, 1], pts[:, 0]] = True
            pg = g.update(); pc = o.update()
            bad = cmp_dm(g.export(O, O, W, W), o.export(O, O, W, W), f"l2={l2} it={it} pops gpu={pg} cpu={pc}") if (pg != pc or it % 10 == 9) else 0
            if pg != pc or bad:
                print("  STOP: divergence", flush=True); break
        pts = np.zeros((500, 3)); pts[:, :2] = rng.uniform(0, W * 0.05, size=(500, 2))
        dg, gg = g.distance(pts); dc, gc = o.distance(pts)
        print(f"  distance max|d| {np.abs(dg-dc).max():.3e} grad {np.abs(gg-gc).max():.3e}", flush=True)


def build_room_dm(l2=1.0):
    ds = synth.make_dataset("loc_room", 3)
    seg = ds.segments
    cells = set()
    for x1, y1, x2, y2 in seg:
        n = int(max(abs(x2 - x1), abs(y2 - y1)) / 0.05) + 1
        for k in range(n + 1):
            x = x1 + (x2 - x1) * k / n; y = y1 + (y2 - y1) * k / n
            cells.add((int(x * 20 + O + 0.5), int(y * 20 + O + 0.5)))
    return ds, np.array(sorted(cells), dtype=np.uint32)


@section("match")
def t_match():
    ds, cells = build_room_dm()
    g = api.DynamicDistanceMap(l2_max=1.0); o = po.DDM(l2_max=1.0)
    g.addObstacle(cells); o.add(cells)
    print("  pops", g.update(), o.update(), flush=True)
    pts = ds.scans[0]
    truth = ds.truth[0]
    states = []
    for dx, dy, dth in ((0.10, -0.07, 0.05), (0, 0, 0), (-0.2, 0.1, -0.08), (0.02, 0.01, 0.3)):
        states.append(po.se2_from_xyr(truth[0] + dx, truth[1] + dy, truth[2] + dth))
    states = np.array(states)
    ne_g = g.matchNormalEquations(pts, states)
    for i, s in enumerate(states):
        ne_c = o.match_normal_eq(pts, s)
        rel = np.abs(ne_g[i, :11] - ne_c) / (np.abs(ne_c) + 1e-30)
        print(f"  normal-eq state {i}: max rel diff {rel.max():.3e}  chi2 {ne_c[9]:.6f}", flush=True)
    for strat in (0, 1):
        sg, stg, sums = g.matchSolve(pts, states, strategy=strat)
        for i, s in enumerate(states):
            sc, _, stc = o.match_solve(pts, s, strategy=strat)
            print(f"  solve strat={strat} state {i}: |dstate| {np.abs(sg[i]-sc).max():.3e} iters gpu {stg[i,0]} cpu {stc[0]} evals gpu {stg[i,1]} cpu {stc[1]} "
                  f"err xy {np.hypot(sc[2]-truth[0], sc[3]-truth[1]):.4f}", flush=True)


@section("loc2d")
def t_loc():
    ds, cells = build_room_dm()
    gl = api.Loc2D(api.Loc2D.Options(trans_thresh=0.01, rot_thresh=0.01))
    ol = po.Loc2D(po.LocOptions.defaults(trans_thresh=0.01, rot_thresh=0.01))
    gl.distance_map.addObstacle(cells); gl.distance_map.update()
    od = ol.dm(); od.add(cells); od.update()
    t0 = ds.truth[0]
    gl.setPose(t0[0] + 0.1, t0[1] - 0.07, t0[2] + 0.05); ol.set_pose(t0[0] + 0.1, t0[1] - 0.07, t0[2] + 0.05)
    for t in range(3):
        a = gl.update(ds.scans[t], ds.odom[t], force_update=(t == 0)); b = ol.update(ds.scans[t], ds.odom[t], force=(t == 0))
        sc, cov, rmse, st = ol.get()
        print(f"  t={t} upd {a}/{b} |dstate| {np.abs(gl.state()-sc).max():.3e} rmse {gl.getRMSE():.6f}/{rmse:.6f} cov rel {np.abs(gl.getCovar()-cov).max()/np.abs(cov).max():.3e} "
              f"stats {gl.solveStats()} {st} err {np.hypot(sc[2]-ds.truth[t,0], sc[3]-ds.truth[t,1]):.4f}", flush=True)


@section("slam2d")
def t_slam():
    for name, T in (("room", 40), ("corridor", 25)):
        ds = synth.make_dataset(name, T)
        g = api.Slam2D(api.Slam2D.Options(trans_thresh=0.05, rot_thresh=0.05))
        o = po.Slam2D(po.SlamOptions.defaults(trans_thresh=0.05, rot_thresh=0.05))
        g.setPose(*ds.truth[0]); o.set_pose(*ds.truth[0])
        for t in range(T):
            a = g.update(ds.scans[t], ds.odom[t]); b = o.update(ds.scans[t], ds.odom[t])
            if t < 3 or t % 10 == 9 or t == T - 1:
                cg, _ = g.counters(); co, _ = o.counters()
                print(f"  {name} t={t} upd {a}/{b} |dstate| {np.abs(g.state()-o.state()).max():.3e} counters gpu {cg} cpu {co}", flush=True)
                n, mn, mx = o.dm_bounds(); w, h = int(mx[0] - mn[0]), int(mx[1] - mn[1])
                cmp_dm(g.exportDistance(int(mn[0]), int(mn[1]), w, h), o.export_dm(mn[0], mn[1], w, h), "dm ")
                n2, mn2, mx2 = o.occ_bounds(); w2, h2 = int(mx2[0] - mn2[0]), int(mx2[1] - mn2[1])
                cmp_occ(g.exportOccupancy(int(mn2[0]), int(mn2[1]), w2, h2), o.export_occ(mn2[0], mn2[1], w2, h2), "occ")
                print(f"  patches gpu occ {g.mapBounds(0)[0]} dm {g.mapBounds(1)[0]} cpu occ {n2} dm {n}", flush=True)


@section("pf")
def t_pf():
    P, T = 16, 60
    ds = synth.make_dataset("room", T)
    g = api.PFSlam2D(api.PFSlam2D.Options(P, trans_thresh=0.05, rot_thresh=0.05, seed=42, timing=1))
    o = po.PFSlam2D(po.PFOptions.defaults(P, trans_thresh=0.05, rot_thresh=0.05, seed=42, threads=8))
    g.setPrior(*ds.truth[0]); o.set_prior(*ds.truth[0])
    for t in range(T):
        a = g.update(ds.scans[t], ds.odom[t]); b = o.update(ds.scans[t], ds.odom[t])
        sg, wg = g.getParticles(); so, wo = o.particles()
        rg, ro = g.lastResample(), o.last_resample()
        same_idx = (len(rg) == len(ro)) and bool((rg == ro).all())
        if t < 3 or t % 10 == 9 or len(ro) or not same_idx:
            cg, _ = g.counters(); co, _ = o.counters()
            print(f"  t={t} upd {a}/{b} |dstate| {np.abs(sg-so).max():.3e} |dw| {np.abs(wg-wo).max():.3e} neff {g.getNeff():.4f}/{o.neff:.4f} "
                  f"resample {len(rg)}/{len(ro)} same {same_idx} best {g.getBestParticleIdx()}/{o.best()}\n     gpu {cg}\n     cpu {co}", flush=True)
    for p in (0, P - 1):
        n, mn, mx = o.dm_bounds(p); w, h = int(mx[0] - mn[0]), int(mx[1] - mn[1])
        cmp_dm(g.exportDistance(p, int(mn[0]), int(mn[1]), w, h), o.export_dm(p, mn[0], mn[1], w, h), f"particle {p} dm ")
        n2, mn2, mx2 = o.occ_bounds(p); w2, h2 = int(mx2[0] - mn2[0]), int(mx2[1] - mn2[1])
        cmp_occ(g.exportOccupancy(p, int(mn2[0]), int(mn2[1]), w2, h2), o.export_occ(p, mn2[0], mn2[1], w2, h2), f"particle {p} occ")
    print("  kernel times", g.kernelTimes(), flush=True)
    print("  traj len", len(g.trajectory(0)), len(o.trajectory(0)), "max diff", np.abs(g.trajectory(0) - o.trajectory(0)).max(), flush=True)


@section("speed")
def t_speed():
    P, T = 256, 30
    ds = synth.make_dataset("room", T)
    g = api.PFSlam2D(api.PFSlam2D.Options(P, trans_thresh=0.05, rot_thresh=0.05, seed=42, timing=1))
    g.setPrior(*ds.truth[0])
    g.update(ds.scans[0], ds.odom[0])
    t0 = time.time()
    for t in range(1, T):
        g.update(ds.scans[t], ds.odom[t])
    dt = time.time() - t0
    print(f"  P={P}: {(T-1)/dt:.1f} scans/s wall; kernel times {g.kernelTimes()} totals {g.counters()[1]}", flush=True)


if __name__ == "__main__":
    print("devices", api.device_count(), api.lib().lama_version().decode(), flush=True)
    which = sys.argv[1:] or ["ddm", "match", "loc", "slam", "pf", "speed"]
    for name, fn in (("ddm", t_ddm), ("match", t_match), ("loc", t_loc), ("slam", t_slam), ("pf", t_pf), ("speed", t_speed)):
        if name in which:
            fn()
