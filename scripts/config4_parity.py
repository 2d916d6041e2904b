"""BASELINE.json configs[3] at its stated size on ONE device: PFSlam2D, 256 particles x 1080 beams, the 5 000-scan synthetic loop, against the
single-process oracle (thread pool on the host cores).  States, weights and resample indices are compared on EVERY scan, the work counters at the
end, every cell of both maps of 8 particles at the end.  Usage: python scripts/config4_parity.py [scans] [threads] [meas_sigma_gain] > profiles/r02_config4_parity.txt
(about 4 minutes of CPU per 5 000 scans on 16 cores)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from iris_lama_b200 import api, synth
from oracle import pyoracle as po

T = int(sys.argv[1]) if len(sys.argv) > 1 else 5001
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 16
gain = float(sys.argv[3]) if len(sys.argv) > 3 else None      # meas_sigma_gain: smaller than the default 0.05 forces resamplings
P, BEAMS = 256, 1080
ds = synth.make_dataset("loop", T, n_beams=BEAMS)
kw = dict(trans_thresh=0.05, rot_thresh=0.05, seed=42)
if gain is not None:
    kw["meas_sigma_gain"] = gain
g = api.PFSlam2D(api.PFSlam2D.Options(P, **kw))
o = po.PFSlam2D(po.PFOptions.defaults(P, threads=threads, **kw))
g.setPrior(*ds.truth[0]); o.set_prior(*ds.truth[0])
worst_state, worst_w, n_res, bad = 0.0, 0.0, 0, 0
tg = to = 0.0
for t in range(T):
    a = time.perf_counter(); ug = g.update(ds.scans[t], ds.odom[t]); sg, wg = g.getParticles(); b = time.perf_counter()
    uo = o.update(ds.scans[t], ds.odom[t]); c = time.perf_counter()
    tg += b - a; to += c - b
    so, wo = o.particles()
    rg, ro = g.lastResample(), o.last_resample()
    n_res += int(len(ro) > 0)
    ds_, dw = float(np.abs(sg - so).max()), float(np.abs(wg - wo).max() / max(1.0, np.abs(wo).max()))
    worst_state, worst_w = max(worst_state, ds_), max(worst_w, dw)
    if ug != uo or rg.tolist() != ro.tolist() or ds_ > 1e-9 or dw > 1e-6:
        bad += 1
        print(f"MISMATCH at scan {t}: updated {ug}/{uo} state diff {ds_:.3e} weight diff {dw:.3e} resample equal {rg.tolist() == ro.tolist()}", flush=True)
        if bad > 5:
            break
    if t % 500 == 0:
        print(f"scan {t}: max state diff so far {worst_state:.3e}, resamplings {n_res}, neff {g.getNeff():.1f}", flush=True)
_, cg = g.counters(); _, co = o.counters()
counters_ok = all(cg[k] == co[k] for k in ("evals", "gn_iters")) and (n_res > 0 or all(cg[k] == co[k] for k in ("ray_cells", "dm_pops")))
cells_ok, cells = True, 0
for p in (0, 31, 64, 101, 128, 177, 200, 255):
    _, mn, mx = o.occ_bounds(p); w, h = int(mx[0] - mn[0]), int(mx[1] - mn[1])
    x, y = g.exportOccupancy(p, int(mn[0]), int(mn[1]), w, h), o.export_occ(p, mn[0], mn[1], w, h)
    cells_ok &= bool((x["occupied"] == y["occupied"]).all() and (x["visited"] == y["visited"]).all() and (x["known"] == y["known"]).all())
    _, mn, mx = o.dm_bounds(p); w, h = int(mx[0] - mn[0]), int(mx[1] - mn[1])
    x, y = g.exportDistance(p, int(mn[0]), int(mn[1]), w, h), o.export_dm(p, mn[0], mn[1], w, h)
    cells_ok &= bool(all((x[k] == y[k]).all() for k in ("sqdist", "valid", "ox", "oy", "queued", "known")))
    cells += 2 * w * h
print(f"config 4 on one device (meas_sigma_gain {gain if gain is not None else 'default'}): {P} particles x {BEAMS} beams x {T - 1} scans after the map-init scan: scans with a mismatch {bad}, max |state diff| {worst_state:.3e}, "
      f"max relative weight diff {worst_w:.3e}, resamplings {n_res} (history digest {g.resampleDigest()}), counters equal {counters_ok} "
      f"(evals {cg['evals']}, ray cells {cg['ray_cells']}, pops {cg['dm_pops']}), cells of 8 particles equal {cells_ok} ({cells} cells compared), "
      f"GPU wall {tg:.1f} s incl. one host sync per scan ({(T - 1) / tg:.0f} scans/s), CPU oracle {to:.1f} s on {threads} threads ({(T - 1) / to:.1f} scans/s)")
print("PARITY", bad == 0 and cells_ok and counters_ok)
