"""Thread scaling of the CPU oracle (thread pool) on this host."""
import sys, time, os
sys.path.insert(0, '.')
from oracle import pyoracle as po
from iris_lama_b200 import synth
P, T = 256, 10
ds = synth.make_dataset("loop", T, n_beams=1080)
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for th in [int(a) for a in sys.argv[1:]] or [1, 8, 32, 64, 128]:
    o = po.PFSlam2D(po.PFOptions.defaults(P, trans_thresh=0.05, rot_thresh=0.05, seed=42, threads=th))
    o.set_prior(*ds.truth[0])
    o.update(ds.scans[0], ds.odom[0]); o.update(ds.scans[1], ds.odom[1])
    t0 = time.perf_counter()
    for t in range(2, T):
        o.update(ds.scans[t], ds.odom[t])
    dt = time.perf_counter() - t0
    print(f"threads {th}: {(T-2)/dt:.2f} scans/s  buckets {o.times()}", flush=True)
