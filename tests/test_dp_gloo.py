"""CPU, world_size 2 over gloo: the host-side data-parallel loop (train.dp_step) that drives the executor's phase API.
A numpy stand-in implements the same phase / sync-point protocol as cenn_trainer_step_phase + cenn_trainer_sync_info
for a miniature "BN + gradient" program; the 2-rank result must equal the single-process result at the global batch."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class FakePhaseTrainer:
    """Program: [sums of x over the local shard] SYNC -> [normalise with the global mean, gradient = sum of xhat * w] SYNC
    -> [loss accumulators (doubles)] SYNC -> update w.  Buffers are numpy arrays addressed by integer 'pointers'."""

    def __init__(self, x_local, n_global):
        self.x, self.n = x_local, n_global
        self.stats = np.zeros(2, np.float32)
        self.grad = np.zeros(3, np.float32)
        self.loss = np.zeros(8, np.float64)
        self.w = np.ones(3, np.float32)
        self.bufs = {1: self.stats, 2: self.grad, 3: self.loss}
        self.phases = [self._p0, self._p1, self._p2, self._p3]
        self.sync = [(1, 2, False), (2, 3, False), (3, 8, True), None]
        self.pc, self.last = 0, None

    def _p0(self):
        self.stats[:] = [self.x.sum(), (self.x ** 2).sum()]

    def _p1(self):
        mean = self.stats[0] / self.n
        var = self.stats[1] / self.n - mean * mean
        xhat = (self.x - mean) / np.sqrt(var + 1e-5)
        self.grad[:] = [xhat.sum(), (xhat ** 2).sum(), (xhat ** 3).sum()]
        self.loss[:] = 0
        self.loss[0] = float((xhat ** 2).sum()) / self.n

    def _p2(self):
        pass

    def _p3(self):
        self.w -= 0.1 * self.grad / self.n

    def step_phase(self, phase, a, b, m):
        if phase < 0:
            self.pc, self.last = 0, None
        while self.pc < len(self.phases):
            self.phases[self.pc]()
            s = self.sync[self.pc]
            self.pc += 1
            if s is not None:
                self.last = s
                return
        self.last = None

    def sync_info(self):
        if self.last is None:
            return None, 0, False, self.pc >= len(self.phases)
        return self.last[0], self.last[1], self.last[2], False


def _worker(rank, world, port, x_full, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from video_filler_b200 import train
    shard = np.array_split(x_full, world)[rank]
    t = FakePhaseTrainer(shard, x_full.size)

    def all_reduce(buf, count, is_double):
        arr = t.bufs[buf]
        assert arr.size == count and (arr.dtype == np.float64) == is_double
        dist.all_reduce(torch.from_numpy(arr))

    n = train.dp_step(t, 0, 0, None, all_reduce)
    out[rank] = (n, t.w.copy(), t.loss.copy(), t.grad.copy())
    dist.destroy_process_group()


def test_dp_step_two_ranks_equal_single_process():
    rng = np.random.default_rng(3)
    x = rng.normal(1.0, 2.0, 64).astype(np.float32)
    from video_filler_b200 import train
    ref = FakePhaseTrainer(x, x.size)
    n_ref = train.dp_step(ref, 0, 0, None, lambda *a: None)
    assert n_ref == 3
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, x, out), nprocs=2, join=True)
    for r in (0, 1):
        n, w, loss, grad = out[r]
        assert n == 3                                        # stats, gradients, loss accumulators
        assert np.allclose(grad, ref.grad, rtol=1e-4, atol=1e-4)
        assert np.allclose(w, ref.w, rtol=1e-5)
        assert loss[0] == pytest.approx(ref.loss[0], rel=1e-5)
    assert np.array_equal(out[0][1], out[1][1])              # replicas stay bit-identical
