"""torchrun --nproc-per-node G scripts/sharded_check.py : sharded PFSlam2D (NCCL) vs the single-process oracle.
Forces resampling (small meas_sigma_gain) so that ancestor maps migrate between ranks."""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, '.')
from iris_lama_b200 import api, synth
from iris_lama_b200.distributed import ShardedPFSlam2D

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
P, T = 16, 30
ds = synth.make_dataset("room", T, n_beams=360)
kw = dict(trans_thresh=0.05, rot_thresh=0.05, seed=5, meas_sigma_gain=0.02)
pf = api.PFSlam2D(api.PFSlam2D.Options(P, device=lr, shard_rank=rank, shard_count=world, **kw))
pf.setPrior(*ds.truth[0])
sh = ShardedPFSlam2D(pf, P, device=dev)
o = None
if rank == 0:
    from oracle import pyoracle as po
    o = po.PFSlam2D(po.PFOptions.defaults(P, **kw)); o.set_prior(*ds.truth[0])
ok = True; n_res = 0
for t in range(T):
    sh.update(ds.scans[t], ds.odom[t])
    if rank == 0:
        o.update(ds.scans[t], ds.odom[t])
        sg, wg = pf.getParticles(); so, wo = o.particles()
        rg, ro = pf.lastResample(), o.last_resample()
        n_res += int(len(ro) > 0)
        good = np.abs(sg - so).max() < 1e-9 and rg.tolist() == ro.tolist() and np.abs(wg - wo).max() < 1e-6 * max(1, np.abs(wo).max())
        ok &= bool(good)
        if not good: print("MISMATCH at", t, np.abs(sg - so).max(), rg.tolist(), ro.tolist(), flush=True)
# map parity of the local particles (rank r checks its own shard against the oracle replayed locally)
from oracle import pyoracle as po2
oo = po2.PFSlam2D(po2.PFOptions.defaults(P, **kw)); oo.set_prior(*ds.truth[0])
for t in range(T): oo.update(ds.scans[t], ds.odom[t])
per = P // world
cells_ok = True
for k in range(per):
    g = rank * per + k
    n, mn, mx = oo.occ_bounds(g); w, h = int(mx[0] - mn[0]), int(mx[1] - mn[1])
    a, b = pf.exportOccupancy(g, int(mn[0]), int(mn[1]), w, h), oo.export_occ(g, mn[0], mn[1], w, h)
    d1, d2 = pf.exportDistance(g, int(mn[0]), int(mn[1]), w, h), oo.export_dm(g, mn[0], mn[1], w, h)
    cells_ok &= bool((a["occupied"] == b["occupied"]).all() and (a["visited"] == b["visited"]).all() and (d1["sqdist"] == d2["sqdist"]).all() and (d1["valid"] == d2["valid"]).all())
flag = torch.tensor([int(ok), int(cells_ok)], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"sharded_check world={world}: trajectory/resample parity {bool(flag[0])}, local map parity on every rank {bool(flag[1])}, resamples {n_res}, "
          f"collectives {sh.collectives}, migrated bytes {sh.migrated_bytes}", flush=True)
dist.destroy_process_group()
sys.exit(0 if (flag[0] and flag[1]) else 1)
