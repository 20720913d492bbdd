"""N>1 path of the scaler action on CPU: two gloo ranks combine FP64 partials with the same
all-reduce code the NCCL path uses (ScalerAccumulator.combine / finalize)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from adyolo_b200.scaler import ScalerAccumulator
    rng = np.random.default_rng(7)
    data = rng.standard_normal((4, 7, 50, 64)) * 10 - 40        # (B, C, T, 64) "features"
    mine = data[rank::world]
    sums = torch.from_numpy(np.stack([mine.sum((0, 2)), (mine ** 2).sum((0, 2))]))
    ext = torch.from_numpy(np.stack([mine.max((0, 2)), mine.min((0, 2))]))
    count, sums, ext = ScalerAccumulator.combine(mine.shape[0] * mine.shape[2], sums, ext)
    mean, std, mx, mn = ScalerAccumulator.finalize(count, sums, ext)
    ret[rank] = (count, mean, std, mx, mn)
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_matches_numpy():
    import adyolo_b200  # noqa: F401  (build check happens in test_cpu_abi)
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    rng = np.random.default_rng(7)
    data = rng.standard_normal((4, 7, 50, 64)) * 10 - 40
    flat = data.transpose(1, 3, 0, 2).reshape(7, 64, -1)
    for r in range(world):
        count, mean, std, mx, mn = ret[r]
        assert count == 200
        np.testing.assert_allclose(mean, flat.mean(-1), rtol=1e-12)
        np.testing.assert_allclose(std, flat.std(-1), rtol=1e-9)
        np.testing.assert_array_equal(mx, flat.max(-1))
        np.testing.assert_array_equal(mn, flat.min(-1))
