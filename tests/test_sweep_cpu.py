"""CPU: the multi-GPU plumbing of sweeps -- contiguous sharding of the batch axis and the gather of the result
spectra -- exercised with world_size 2 and 3 on the gloo backend (the N>1 path of bench.py / sweep.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyhalomodel_b200 import sweep


def test_shard_bounds_cover_everything_once():
    for n in (1, 2, 7, 8, 4096, 16385):
        for world in (1, 2, 3, 4, 8):
            b = sweep.shard_bounds(n, world)
            assert len(b) == world and b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i+1][0] for i in range(world-1))
            sizes = [hi-lo for lo, hi in b]
            assert max(sizes)-min(sizes) <= 1 and sum(sizes) == n


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, result_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        lo, hi = sweep.shard_range(n_items, rank, world)
        S, nk = 3, 5
        # every "spectrum" encodes its global model index, so ordering mistakes are visible
        idx = torch.arange(lo, hi, dtype=torch.float64)
        local = idx[:, None, None, None]*1000.+torch.arange(S*3*nk, dtype=torch.float64).reshape(S, 3, nk)
        full = sweep.gather_spectra(local, n_items)
        want = torch.arange(n_items, dtype=torch.float64)[:, None, None, None]*1000. \
            + torch.arange(S*3*nk, dtype=torch.float64).reshape(S, 3, nk)
        assert full.shape == want.shape and torch.equal(full, want)

        class FakePlan:
            def __init__(self, lo, hi, F):
                self.out = (torch.arange(lo*F, hi*F, dtype=torch.float64)[:, None, None, None]
                            + torch.zeros(S, 3, nk, dtype=torch.float64))

            def run(self):
                return self.out, None

        class FakeBatch:
            def __init__(self, lo, hi):
                self.lo, self.hi = lo, hi

            def plan(self):
                return FakePlan(self.lo, self.hi, 5)

        sw = sweep.ShardedSweep(n_items, FakeBatch, models_per_sample=5, tail_shape=(S, 3, nk))
        out = sw.run()
        assert out.shape == (n_items*5, S, 3, nk)
        assert torch.equal(out[:, 0, 0, 0], torch.arange(n_items*5, dtype=torch.float64))
        np.save(os.path.join(result_dir, 'ok_%d.npy' % rank), np.array([1]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,n_items', [(2, 8), (2, 7), (3, 10), (3, 2)])     # (3, 2): one rank owns no sample
def test_gather_over_gloo(tmp_path, world, n_items):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_items, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path/('ok_%d.npy' % r)) for r in range(world))


def test_gather_without_process_group_is_identity():
    x = torch.arange(12.).reshape(4, 3)
    assert sweep.gather_spectra(x, 4) is x
