"""torchrun --nproc-per-node G scripts/sharded_native_check.py [P T beams gain] : the sharded step INSIDE the library (lama_pf_shard_connect +
lama_pf_update: NCCL all-gather of the match results, NCCL send / recv of migrating maps, all from C++) against the single-process
oracle.  Forces resampling (small meas_sigma_gain) so that ancestor maps migrate between ranks.  torch.distributed only hands the
128-byte NCCL id from rank 0 to the others and reduces the verdict."""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, '.')
from iris_lama_b200 import api, synth

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(lr)
dist.init_process_group("gloo")
P = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 30
beams = int(sys.argv[3]) if len(sys.argv) > 3 else 360
gain = float(sys.argv[4]) if len(sys.argv) > 4 else 0.02
world_name = "room" if beams <= 720 else "loop"
ds = synth.make_dataset(world_name, T, n_beams=beams)
kw = dict(trans_thresh=0.05, rot_thresh=0.05, seed=5, meas_sigma_gain=gain)
pf = api.PFSlam2D(api.PFSlam2D.Options(P, device=lr, shard_rank=rank, shard_count=world, **kw))
pf.setPrior(*ds.truth[0])
box = [api.shard_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
pf.shardConnect(box[0])
from oracle import pyoracle as po
o = po.PFSlam2D(po.PFOptions.defaults(P, threads=8, **kw)); o.set_prior(*ds.truth[0])
ok = True; n_res = 0
for t in range(T):
    a = pf.update(ds.scans[t], ds.odom[t])
    b = o.update(ds.scans[t], ds.odom[t])
    sg, wg = pf.getParticles(); so, wo = o.particles()
    rg, ro = pf.lastResample(), o.last_resample()
    n_res += int(len(ro) > 0)
    good = a == b and np.abs(sg - so).max() < 1e-9 and rg.tolist() == ro.tolist() and np.abs(wg - wo).max() < 1e-6 * max(1, np.abs(wo).max())
    ok &= bool(good)
    if not good: print(f"rank {rank} MISMATCH at", t, np.abs(sg - so).max(), rg.tolist()[:8], ro.tolist()[:8], flush=True)
per = P // world
cells_ok = True
for k in range(per):
    g = rank * per + k
    n, mn, mx = o.occ_bounds(g); w, h = int(mx[0] - mn[0]), int(mx[1] - mn[1])
    a, b = pf.exportOccupancy(g, int(mn[0]), int(mn[1]), w, h), o.export_occ(g, mn[0], mn[1], w, h)
    n, mn, mx = o.dm_bounds(g); w, h = int(mx[0] - mn[0]), int(mx[1] - mn[1])
    d1, d2 = pf.exportDistance(g, int(mn[0]), int(mn[1]), w, h), o.export_dm(g, mn[0], mn[1], w, h)
    cells_ok &= bool((a["occupied"] == b["occupied"]).all() and (a["visited"] == b["visited"]).all() and all((d1[f] == d2[f]).all() for f in ("sqdist", "valid", "ox", "oy", "queued")))
coll, mig = pf.shardStats()
flag = torch.tensor([int(ok), int(cells_ok), -mig]); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"sharded_native_check world={world} P={P} beams={beams} T={T}: states/weights/resample indices on every rank {bool(flag[0])}, local map cells on every rank "
          f"{bool(flag[1])}, resamples {n_res}, collectives (rank 0) {coll}, most bytes of maps received by one rank {-int(flag[2])}", flush=True)
dist.destroy_process_group()
sys.exit(0 if (flag[0] and flag[1]) else 1)
