import sys, time, numpy as np
sys.path.insert(0, '.')
from oracle import pyoracle as po
from iris_lama_b200 import synth

T = int(sys.argv[1]) if len(sys.argv) > 1 else 100
P = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ds = synth.make_dataset("room", T)
print("dataset", ds.scans.shape)
opts = po.PFOptions.defaults(P, trans_thresh=0.05, rot_thresh=0.05, seed=42, threads=8)
pf = po.PFSlam2D(opts)
pf.set_prior(*ds.truth[0])
t0 = time.time()
for t in range(T):
    ok = pf.update(ds.scans[t], ds.odom[t])
    last, tot = pf.counters()
    if t < 5 or t % 20 == 0:
        st, w = pf.particles()
        b = pf.best()
        print(t, ok, last, "neff %.2f" % pf.neff, "best", b, "err", np.hypot(st[b,2]-ds.truth[t,0], st[b,3]-ds.truth[t,1]))
print("time", time.time()-t0, pf.times(), tot)
