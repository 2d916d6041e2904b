"""world_size-2 gloo test of the particle-sharding orchestration (iris_lama_b200/distributed.py): all-gather of match
results (+ the digest of the previous resampling decision), point-to-point migration of ancestor maps.  The device object is replaced by
a stub that implements the same shard*/pack/unpack calls, so the host-side plan is tested without a GPU."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from iris_lama_b200.distributed import ShardedPFSlam2D, local_sources, migration_plan  # noqa: E402


class StubPF:
    """Same call protocol as api.PFSlam2D in sharded mode.  A particle's "map" is the list of scan ids it absorbed,
    prefixed by its lineage, so migrations are observable."""

    def __init__(self, P, rank, world, seed=0):
        self.P, self.rank, self.world, self.per = P, rank, world, P // world
        self.rng = np.random.default_rng(seed)          # same seed on every rank, like the host mt19937
        self.maps = [[] for _ in range(2 * self.per)]   # local slots + staging slots
        self.t = 0
        self.first = True
        self.weights = np.zeros(P)

    def shardBegin(self, pts, odom, timestamp=0.0):
        self.t += 1
        noise = self.rng.normal(size=(self.P, 5))       # every rank draws for ALL particles
        if self.first:
            self.first = False
            for k in range(self.per):
                self.maps[k] = [("init", 0)]
            return 1, np.zeros((self.per, 5))
        lo = self.rank * self.per
        return 2, noise[lo:lo + self.per].copy()

    def shardFinish(self, allr):
        assert allr.shape == (self.P, 5)
        self.weights = allr[:, 4]
        resample = (self.t % 3 == 0)
        u = self.rng.random()                            # consumed on every rank
        idx = np.sort(self.rng.integers(0, self.P, self.P)).astype(np.int32) if resample else np.zeros(self.P, np.int32)
        return resample, idx

    def packParticle(self, slot):
        flat = np.array([v for _, v in self.maps[slot]], np.int64)
        return flat.view(np.uint8).copy()

    def unpackParticle(self, slot, buf):
        vals = np.frombuffer(np.ascontiguousarray(buf).tobytes(), np.int64)
        self.maps[slot] = [("mig", int(v)) for v in vals]

    def shardApply(self, idx, local_src=None):
        new = [list(self.maps[int(s)]) for s in local_src]
        for k in range(self.per):
            self.maps[k] = new[k]
        for k in range(self.per, 2 * self.per):
            self.maps[k] = []

    def shardMapUpdate(self):
        lo = self.rank * self.per
        for k in range(self.per):
            self.maps[k].append(("scan", 1000 * self.t + lo + k))


def _worker(rank, world, port, P, T, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pf = StubPF(P, rank, world)
    sh = ShardedPFSlam2D(pf, P)
    for t in range(T):
        sh.update(np.zeros((4, 3)), np.zeros(3))
    out[rank] = [[v for _, v in m] for m in pf.maps[:pf.per]]
    dist.destroy_process_group()


def _single_process_truth(P, T):
    pf = StubPF(P, 0, 1)
    for t in range(T):
        did, local = pf.shardBegin(None, None)
        if did == 2:
            res, idx = pf.shardFinish(local)
            if res:
                pf.shardApply(idx, idx)
            pf.shardMapUpdate()
    return [[v for _, v in m] for m in pf.maps[:P]]


def test_migration_plan_and_local_sources():
    idx = np.array([0, 0, 5, 5, 5, 6, 7, 7], np.int32)
    need, serve = migration_plan(idx, 2)
    assert need == [[5], []]              # rank 0 (particles 0-3) needs 5; rank 1's ancestors 5,6,7 are local
    assert serve == [[], [(0, 5)]]
    assert local_sources(idx, 0, 2, need[0]).tolist() == [0, 0, 4, 4]   # staging slot P_local + 0
    assert local_sources(idx, 1, 2, need[1]).tolist() == [1, 2, 3, 3]
    idx = np.array([4, 4, 4, 4, 0, 0, 1, 2], np.int32)
    need, serve = migration_plan(idx, 2)
    assert need == [[4], [0, 1, 2]] and serve == [[(1, 0), (1, 1), (1, 2)], [(0, 4)]]


@pytest.mark.parametrize("world", [2])
def test_sharded_update_equals_single_process(world):
    P, T = 8, 10
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, P, T, out), nprocs=world, join=True)
    truth = _single_process_truth(P, T)
    per = P // world
    for r in range(world):
        for k in range(per):
            got = out[r][k]
            want = truth[r * per + k]
            # the per-scan ids encode which global particle slot absorbed the scan: identical lineages required
            assert got == want, (r, k, got, want)


@pytest.mark.parametrize("G", [2, 4])
def test_logical_ranks_in_one_process_equal_single_process(G):
    """distributed.LocalShards (what the GPU suite uses for its G-logical-ranks test) drives the same shard protocol"""
    from iris_lama_b200.distributed import LocalShards
    P, T = 8, 10
    hs = [StubPF(P, r, G) for r in range(G)]
    sh = LocalShards(hs, P)
    for t in range(T):
        sh.update(np.zeros((4, 3)), np.zeros(3))
    truth = _single_process_truth(P, T)
    per = P // G
    for r in range(G):
        for k in range(per):
            assert [v for _, v in hs[r].maps[k]] == truth[r * per + k], (r, k)
    assert sh.migrated_bytes > 0
