"""torchrun --nproc-per-node G scripts/sharded_phases.py : host-side time of every phase of a sharded step (256 x 1080, steady regime)."""
import os, sys, time
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, '.')
from iris_lama_b200 import api, synth
from iris_lama_b200.distributed import ShardedPFSlam2D

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
P, T0, T1 = 256, 200, 260
ds = synth.make_dataset("loop", T1, n_beams=1080)
pf = api.PFSlam2D(api.PFSlam2D.Options(P, device=lr, shard_rank=rank, shard_count=world, trans_thresh=0.05, rot_thresh=0.05, seed=42))
pf.setPrior(*ds.truth[0])
sh = ShardedPFSlam2D(pf, P, device=dev)
acc = {}
def tick(name, t0):
    t1 = time.perf_counter(); acc[name] = acc.get(name, 0.0) + (t1 - t0); return t1
for t in range(T1):
    if t < T0:
        sh.update(ds.scans[t], ds.odom[t]); continue
    if t == T0:
        torch.cuda.synchronize(); dist.barrier(); t_all = time.perf_counter()
    t0 = time.perf_counter()
    did, local = pf.shardBegin(ds.scans[t], ds.odom[t], 0.0); t0 = tick("shardBegin (sampling, enqueue, wait for match)", t0)
    with sh._comm():
        payload = np.concatenate([local.reshape(-1), [0.0]])
        mine = sh._t(payload, torch.float64)
        allr = torch.empty(world * (P // world * 5 + 1), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, mine)
        g = allr.cpu().numpy().reshape(world, P // world * 5 + 1)
    all_results = np.ascontiguousarray(g[:, :-1]).reshape(P, 5)
    t0 = tick("all_gather incl. copies", t0)
    resampled, idx = pf.shardFinish(all_results); t0 = tick("shardFinish (normalise, resampling decision)", t0)
    pf.shardMapUpdate(); t0 = tick("shardMapUpdate", t0)
torch.cuda.synchronize(); dist.barrier()
tot = time.perf_counter() - t_all
if rank == 0:
    n = T1 - T0
    print(f"world={world}: {n / tot:.1f} scans/s, {1e3 * tot / n:.3f} ms per step")
    for k, v in acc.items(): print(f"  {k}: {1e3 * v / n:.3f} ms")
dist.destroy_process_group()
