"""Multi-GPU paths on real GPUs (skipped on a 1-GPU box): the scaler action sharded over 2 ranks
with an NCCL all-reduce must equal the single-process result; the clip-sharded front end needs no
collective at all (each rank's features equal the single-GPU features of its clips)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, clips, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import adyolo_b200 as A
    res = A.preprocess_scaler(clips, batch_clips=2)
    feats = A.features_batched(torch.from_numpy(np.stack(clips[rank::world])).cuda()).cpu().numpy()
    ret[rank] = (res, feats)
    dist.destroy_process_group()


def test_scaler_two_ranks_nccl_matches_single_process():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import __graft_entry__ as g
    g.build()
    import adyolo_b200 as A
    import torch.multiprocessing as mp
    rng = np.random.default_rng(4)
    clips = [np.clip(rng.standard_normal((24000 * 2, 4)) * (500 + 800 * i), -32768, 32767).astype(np.int16) for i in range(6)]
    clips[2][10000:30000] = 0
    single = A.preprocess_scaler(clips, batch_clips=2, rank=0, world_size=1)
    feats_all = A.features_batched(torch.from_numpy(np.stack(clips)).cuda()).cpu().numpy()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29500 + os.getpid() % 2000, clips, ret), nprocs=2, join=True)
    for r in range(2):
        res, feats = ret[r]
        for grp in ("MEL", "IV"):
            for k in ("mean", "std"):
                np.testing.assert_allclose(res[grp][k], single[grp][k], rtol=1e-12, atol=1e-14)
            np.testing.assert_array_equal(res[grp]["max"], single[grp]["max"])
            np.testing.assert_array_equal(res[grp]["min"], single[grp]["min"])
        np.testing.assert_array_equal(feats, feats_all[r::2])
