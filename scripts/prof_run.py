"""Short PFSlam2D run at the benchmark size for ncu captures (launch list / --set full)."""
import sys
sys.path.insert(0, '.')
from iris_lama_b200 import api, synth
P = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 14
ds = synth.make_dataset("loop", T, n_beams=1080)
g = api.PFSlam2D(api.PFSlam2D.Options(P, trans_thresh=0.05, rot_thresh=0.05, seed=42))
g.setPrior(*ds.truth[0])
for t in range(T):
    g.update(ds.scans[t], ds.odom[t])
print(g.counters()[1])
