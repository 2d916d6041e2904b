"""Developer experiment: ray-cast stage time (k_ray_setup + k_ray_pull or k_raycast) at several particle counts, kernel events of the non-pipelined timing mode.
usage: LAMA_B200_LIB=... LAMA_PULL_MAX_PARTICLES=... python scripts/pull_variants.py 32 256"""
import os, sys
sys.path.insert(0, '.')
from iris_lama_b200 import api, synth
T = 360
ds = synth.make_dataset("loop", T, n_beams=1080)
for P in [int(a) for a in sys.argv[1:]] or [32]:
    g = api.PFSlam2D(api.PFSlam2D.Options(P, trans_thresh=0.05, rot_thresh=0.05, seed=42, timing=1))
    g.setPrior(*ds.truth[0])
    for t in range(300):
        g.update(ds.scans[t], ds.odom[t])
    g.getPose(); a, _ = g.kernelTimes()
    for t in range(300, T):
        g.update(ds.scans[t], ds.odom[t])
    g.getPose(); b, _ = g.kernelTimes()
    print(os.environ.get("LAMA_B200_LIB", "default").split("/")[-1], "pull_max", os.environ.get("LAMA_PULL_MAX_PARTICLES", "48"), "P", P,
          {k: round((b[k] - a[k]) / (T - 300), 4) for k in a}, flush=True)
    del g
