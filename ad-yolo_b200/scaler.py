"""`scaler` action: per-(mel bin, channel) mean/std/max/min of un-standardised features.

Reference: ``preprocess.py:87-130`` concatenates the features of every training file in host RAM
(O(dataset), ~310 GB for 600 h) and calls numpy.  Here clips are streamed through the fused
front end in batches; FP64 {count, sum, sum-of-squares} and {max, min} partials live on the device
and, when ``torch.distributed`` is initialised (one process per GPU, clips sharded by rank), are
combined with one SUM and one MAX all-reduce (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import ctypes as C
import os
import pickle

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, require_cuda, stream_ptr
from .features import features_batched, features_mic_batched


class ScalerAccumulator:
    """Streaming statistics over feature batches (B, Cf, T, 64)."""

    def __init__(self, n_channels: int = 7, device="cuda"):
        self.C = n_channels
        self.device = torch.device(device)
        self.count = 0
        if self.device.type == "cuda":
            self._init_buffers()

    def _init_buffers(self):
        self.sums = torch.zeros((2, self.C, 64), dtype=torch.float64, device=self.device)        # sum, sumsq
        self.ext = torch.empty((2, self.C, 64), dtype=torch.float64, device=self.device)         # max, -min later
        self.ext[0].fill_(-float("inf"))
        self.ext[1].fill_(float("inf"))

    def update(self, feats: torch.Tensor):
        """feats: un-standardised, top_db-clamped features (B, C, T, 64) float32 on CUDA."""
        require_cuda(feats, "ScalerAccumulator.update")
        feats = feats.contiguous()
        B, Cf, T, F = feats.shape
        if Cf != self.C or F != 64:
            raise ValueError("feature batch must be (B, %d, T, 64)" % self.C)
        with torch.cuda.device(feats.device):
            check(_lib.lib().adyolo_scaler_partials(ptr(feats), B, Cf, T, ptr(self.sums[0]), ptr(self.sums[1]),
                                                    ptr(self.ext[0]), ptr(self.ext[1]), stream_ptr()),
                  "adyolo_scaler_partials")
        self.count += B * T

    def update_from_audio(self, audio_i16: torch.Tensor):
        """int16 clips (B, N, 4) on the device -> statistics of their un-standardised features (top_db clamp per
        clip = per file, preprocess.py:111).  FOA when the accumulator has 7 channels, MIC (GCC-PHAT) for 10."""
        feats = features_batched(audio_i16, None) if self.C == 7 else features_mic_batched(audio_i16, None)
        self.update(feats)

    # ---- combination across ranks + finalisation (pure tensor plumbing; also used by gloo tests)
    @staticmethod
    def combine(count: int, sums: torch.Tensor, ext: torch.Tensor):
        """All-reduce partials over the default process group (no-op when not initialised): one SUM over
        {sum, sumsq, count} and one MAX over {max, -min}, both enqueued on the current stream right after the
        partials kernels (torch.distributed's NCCL backend is stream-ordered: no host round trip here; the
        frame count stays a device scalar until ``finalize`` reads the results back)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            packed = torch.cat([sums.reshape(-1), torch.full((1,), float(count), dtype=torch.float64, device=sums.device)])
            dist.all_reduce(packed, op=dist.ReduceOp.SUM)
            mx = torch.stack([ext[0], -ext[1]])              # max and max(-x) == -min
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sums = packed[:-1].reshape(sums.shape)
            count = packed[-1]
            ext = torch.stack([mx[0], -mx[1]])
        return count, sums, ext

    @staticmethod
    def finalize(count: int, sums: torch.Tensor, ext: torch.Tensor):
        """-> mean, std (ddof=0), max, min as (C, 64) float64 numpy arrays."""
        n = float(count)                                     # (device scalar after an all-reduce: the one sync)
        if n <= 0:
            raise ValueError("scaler statistics over zero frames (empty shard and no all-reduce?)")
        mean = sums[0] / n
        var = torch.clamp(sums[1] / n - mean * mean, min=0.0)
        return (mean.cpu().numpy(), torch.sqrt(var).cpu().numpy(), ext[0].cpu().numpy(), ext[1].cpu().numpy())

    def result(self, groups=(("MEL", 4), ("IV", 3))):
        """The reference's pickle schema: {'MEL': {mean,std,max,min}, 'IV': {...}} with (1,64,C) arrays."""
        count, sums, ext = self.combine(self.count, self.sums, self.ext)
        mean, std, mx, mn = self.finalize(count, sums, ext)
        out, c0 = {}, 0
        for name, nc in groups:
            sl = slice(c0, c0 + nc)
            out[name] = {"mean": mean[sl].T[None].copy(), "std": std[sl].T[None].copy(),
                         "max": mx[sl].T[None].copy(), "min": mn[sl].T[None].copy()}
            c0 += nc
        return out


def preprocess_scaler(clips, fmt: str = "foa", batch_clips: int = 8, out_path: str | None = None,
                      rank: int | None = None, world_size: int | None = None):
    """preprocess.py:87-130 over an iterable / list of int16 (N, 4) clips (equal length per batch).

    With torch.distributed initialised, each rank processes ``clips[rank::world_size]`` and the
    statistics are all-reduced; every rank returns the same dict (rank 0 writes ``out_path``).
    The top_db clamp unit is the clip (the file, preprocess.py:111).
    """
    import torch.distributed as dist
    require_cuda(None, "preprocess_scaler")
    if (rank is None) != (world_size is None):
        raise ValueError("preprocess_scaler: pass rank and world_size together (or neither: both come from torch.distributed)")
    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if not 0 <= rank < world_size:
        raise ValueError(f"preprocess_scaler: rank {rank} outside world_size {world_size}")
    nC = 7 if fmt == "foa" else 10
    acc = ScalerAccumulator(nC, "cuda")
    mine = list(clips)[rank::world_size]
    i = 0
    while i < len(mine):
        n0 = len(mine[i])
        j = i
        while j < len(mine) and j - i < batch_clips and len(mine[j]) == n0:
            j += 1
        batch = torch.from_numpy(np.stack([np.asarray(c, dtype=np.int16) for c in mine[i:j]])).cuda()
        feats = features_batched(batch, None) if fmt == "foa" else features_mic_batched(batch, None)
        acc.update(feats)
        i = j
    res = acc.result((("MEL", 4), ("IV", 3)) if fmt == "foa" else (("MEL", 4), ("GCC", 6)))
    if out_path and rank == 0:
        with open(out_path, "wb") as f:
            pickle.dump(res, f)
    return res
