"""Developer experiment (make EXTRA=-DLAMA_PHASE_TIMING): per-particle phase cycles of k_brushfire in the revisit regime (scan 330) and while exploring (scan 20)."""
import os, sys
sys.path.insert(0, '.')
which = sys.argv[1] if len(sys.argv) > 1 else "330"
os.environ["LAMA_BF_DEBUG"] = which
from iris_lama_b200 import api, synth
n = int(which) + 3
ds = synth.make_dataset("loop", n, n_beams=1080)
g = api.PFSlam2D(api.PFSlam2D.Options(256, trans_thresh=0.05, rot_thresh=0.05, seed=42))
g.setPrior(*ds.truth[0])
for t in range(n):
    g.update(ds.scans[t], ds.odom[t])
g.getPose()
