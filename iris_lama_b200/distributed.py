"""Multi-GPU particle sharding: one process per GPU, torch.distributed (NCCL) for the plumbing.

Particles shard embarrassingly: particle i lives on rank i // (P / G).  Per scan the only
exchange is (1) an all-gather of the P x 5 match results (SE2 state + log-likelihood, 10 KB at P = 256) and
(2) a broadcast of the P resampling indices; on resampling scans the maps of offspring whose ancestor lives on
another rank migrate point-to-point.  The reference has no distributed layer at all (its only parallelism is the
thread pool of src/pf_slam2d.cpp:254-266,292-302).

`ShardedPFSlam2D` works with any object implementing the shard* / pack / unpack calls of api.PFSlam2D, which is
how the gloo CPU tests exercise the orchestration without a GPU.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def migration_plan(idx: np.ndarray, n_ranks: int):
    """For every rank, which remote ancestors it needs and where they come from.

    Returns (need, serve): need[r] = sorted unique global ancestor ids rank r must fetch,
    serve[r] = list of (dst_rank, global_id) rank r must send, in a deterministic order."""
    P = len(idx)
    per = P // n_ranks
    need = []
    for r in range(n_ranks):
        anc = np.unique(idx[r * per:(r + 1) * per])
        need.append([int(a) for a in anc if a // per != r])
    serve = [[] for _ in range(n_ranks)]
    for r in range(n_ranks):
        for a in need[r]:
            serve[a // per].append((r, a))
    return need, serve


def local_sources(idx: np.ndarray, rank: int, n_ranks: int, need_r):
    """Engine slot every new local particle copies from: local ancestors map to their own slot, remote ones
    to the staging slot (P_local + position in need_r) they were unpacked into."""
    P = len(idx)
    per = P // n_ranks
    stage = {a: per + k for k, a in enumerate(need_r)}
    out = np.zeros(per, np.int32)
    for k in range(per):
        a = int(idx[rank * per + k])
        out[k] = a - rank * per if a // per == rank else stage[a]
    return out


class ShardedPFSlam2D:
    def __init__(self, pf, particles: int, device=None, group=None):
        self.pf = pf
        self.P = particles
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.per = particles // self.world
        self.device = device if device is not None else torch.device("cpu")
        self.collectives = 0
        self.migrated_bytes = 0
        # The collectives get their own CUDA stream: NCCL orders itself after the work already enqueued on the stream it is
        # called from, and the engine's stream holds this scan's map update (enqueued together with the match), which the
        # exchange of the match results must not wait for.
        self.comm_stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        self._digest = 0.0   # of the previous scan's resampling decision (see update)

    def _comm(self):
        import contextlib
        return torch.cuda.stream(self.comm_stream) if self.comm_stream is not None else contextlib.nullcontext()

    def _t(self, a, dtype):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dtype).to(self.device)

    def update(self, pts, odom, timestamp=0.0) -> bool:
        did, local = self.pf.shardBegin(pts, odom, timestamp)
        if did != 2:
            return did != 0
        # ONE collective per scan: all-gather of the local match results.  Every rank then normalises and resamples on the same
        # gathered bytes with the same code and RNG state, so all ranks take the same decision; as a cross-check each rank appends
        # a digest of its PREVIOUS decision to its payload and everybody compares the gathered digests (a divergence is reported
        # one scan late instead of costing a second collective on every scan).
        with self._comm():
            payload = np.concatenate([local.reshape(-1), [self._digest]])
            mine = self._t(payload, torch.float64)
            allr = torch.empty(self.world * (self.per * 5 + 1), dtype=torch.float64, device=self.device)
            dist.all_gather_into_tensor(allr, mine, group=self.group)
            self.collectives += 1
            g = allr.cpu().numpy().reshape(self.world, self.per * 5 + 1)
        if not (g[:, -1] == g[0, -1]).all():
            raise RuntimeError("resampling decision diverged between ranks")
        all_results = np.ascontiguousarray(g[:, :-1]).reshape(self.P, 5)
        resampled, idx = self.pf.shardFinish(all_results)
        self._digest = float(int(resampled) + (int(np.dot(idx.astype(np.int64), np.arange(1, self.P + 1, dtype=np.int64)) % 9007199254740881) if resampled else 0))
        if resampled:
            self._migrate_and_apply(idx)
        self.pf.shardMapUpdate()
        return True

    def _migrate_and_apply(self, idx):
        with self._comm():
            self._migrate_and_apply_impl(idx)

    def _migrate_and_apply_impl(self, idx):
        need, serve = migration_plan(idx, self.world)
        # sizes first (all ranks know who sends what; only byte counts are unknown)
        bufs = [(dst, gid, self.pf.packParticle(gid - self.rank * self.per)) for dst, gid in serve[self.rank]]
        sizes = torch.zeros(self.world, self.P, dtype=torch.int64, device=self.device)
        for dst, gid, b in bufs:
            sizes[dst, gid] = b.size
        dist.all_reduce(sizes, group=self.group)
        self.collectives += 1
        sizes = sizes.cpu().numpy()
        ops, recv = [], []
        for dst, gid, b in bufs:
            t = self._t(b, torch.uint8)
            ops.append(dist.P2POp(dist.isend, t, dst, group=self.group))
        for a in need[self.rank]:
            t = torch.empty(int(sizes[self.rank, a]), dtype=torch.uint8, device=self.device)
            recv.append((a, t))
            ops.append(dist.P2POp(dist.irecv, t, a // self.per, group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for k, (a, t) in enumerate(recv):
            self.pf.unpackParticle(self.per + k, t.cpu().numpy())
            self.migrated_bytes += t.numel()
        self.pf.shardApply(idx, local_sources(idx, self.rank, self.world, need[self.rank]))


class LocalShards:
    """G logical ranks in ONE process (all handles on the same device): the orchestration of ShardedPFSlam2D with the exchange done
    by plain copies instead of collectives.  It drives exactly the calls a rank makes (shardBegin / shardFinish / packParticle /
    unpackParticle / shardApply with local sources / shardMapUpdate), so that the sharded code path can be checked against the
    single-process oracle where only one GPU (or none: any object with the shard* calls works) is available."""

    def __init__(self, handles, particles: int):
        self.h = list(handles)
        self.G = len(self.h)
        self.P = particles
        self.per = particles // self.G
        self.migrated_bytes = 0
        self.last_idx = np.zeros(0, np.int32)

    def update(self, pts, odom, timestamp=0.0) -> bool:
        begun = [h.shardBegin(pts, odom, timestamp) for h in self.h]
        dids = {d for d, _ in begun}
        assert len(dids) == 1, "the ranks disagree about the motion gate"
        did = dids.pop()
        self.last_idx = np.zeros(0, np.int32)
        if did != 2:
            return did != 0
        all_results = np.concatenate([loc for _, loc in begun], axis=0).reshape(self.P, 5)   # the all-gather
        fin = [h.shardFinish(all_results) for h in self.h]
        resampled = {r for r, _ in fin}
        assert len(resampled) == 1, "the ranks disagree about resampling"
        if resampled.pop():
            idx = fin[0][1]
            for r, i in fin[1:]:
                assert (i == idx).all(), "the ranks drew different resampling indices"
            self.last_idx = idx.copy()
            need, serve = migration_plan(idx, self.G)
            packed = {}
            for r in range(self.G):   # every rank packs what it serves ...
                for dst, gid in serve[r]:
                    packed[(dst, gid)] = self.h[r].packParticle(gid - r * self.per)
            for r in range(self.G):   # ... and unpacks what it needs into its staging slots
                for k, a in enumerate(need[r]):
                    buf = packed[(r, a)]
                    self.h[r].unpackParticle(self.per + k, buf)
                    self.migrated_bytes += buf.size
                self.h[r].shardApply(idx, local_sources(idx, r, self.G, need[r]))
        for h in self.h:
            h.shardMapUpdate()
        return True
