"""SpecAug on the device: mirror of ``utils/augmentations.py:6-33``.

The reference applies ``SpecAug.augment`` in ``Dataset.__getitem__`` (``datasets.py:158-160``) to
each feature group (MEL, IV) permuted to (C, T, F).  torchaudio masks the last two axes of whatever
it is given, so with that layout ``TimeMasking`` zeroes a band of **mel bins** (width drawn from
``spec_augment_time_mask_param``) and ``FrequencyMasking`` a run of **frames**
(``spec_augment_freq_mask_param``); every channel of the group shares the mask and the groups are
drawn independently.  This module keeps that behaviour:

* the intervals are drawn on the host with the reference's RNG sequence -- ``random.random()`` for
  each of the two gates, then ``torch.rand(1)`` twice per applied mask with torchaudio 2.x's
  float32 arithmetic -- so the same python/torch seeds give the same masks as the reference class
  (pinned by ``tests/golden/specaug.npz``);
* ``adyolo_spec_mask`` writes the zeros into the (B, C, T, F) tensor the fused front end produced
  (write-only kernel, one launch per batch).

Rotation augmentation (the other half of ``augmentations.py``) is fused into the front end and the
label kernel: pass ``rot_comb`` to ``features_batched`` / ``label_rows_batched``.
"""
from __future__ import annotations

import ctypes as C
import random

import torch

from . import _lib
from ._lib import check, ptr, require_cuda, stream_ptr

FOA_GROUPS = ((0, 4), (4, 7))       # [mel W,Y,Z,X | iv Y,Z,X]   datasets.py:158-161
# RotationAug combination number (augmentations.py:46-70) -> bit0 = Y negated, bit1 = Z negated, bit2 = X negated,
# bit3 = X <-> Y swap (the same table the kernels use, csrc/fe2_core.cuh::rot_bits_rt)
ROT_BITS = (0x0, 0x2, 0x1, 0x3, 0x5, 0x7, 0x4, 0x6, 0x9, 0xB, 0x8, 0xA, 0xC, 0xE, 0xD, 0xF)
MIC_GROUPS = ((0, 4), (4, 10))      # [mel x4 | gcc x6]


class RotationAug(object):
    """Host half of ``augmentations.py:36-111``: same constructor and on/off rule; ``draw`` picks one of
    the 16 channel-sign / swap combinations per clip with the reference's own call,
    ``int(random.uniform(0, 16))`` (``:93``), and returns them as the int8 device tensor that
    ``features_batched(..., rot_comb=)`` and ``label_rows_batched(..., rot_comb=)`` take -- the
    rotation itself happens inside those kernels (audio channel signs / X<->Y swap folded into the
    front end, azimuth / elevation transform into the label kernel)."""

    def __init__(self, params: dict, is_valid: bool):
        self.apply_augment = bool(params["aug_config"]["rotation_augment"]) and not is_valid

    def draw(self, n_clips: int, device) -> torch.Tensor | None:
        """int8 (n_clips,) combination numbers on ``device``; ``None`` when augmentation is off."""
        if not self.apply_augment:
            return None
        comb = [int(random.uniform(0, 16)) for _ in range(n_clips)]
        return torch.tensor(comb, dtype=torch.int8).to(device, non_blocking=True)


def rotate_host(audio_i16, events, comb_no: int):
    """``RotationAug._rotate`` (augmentations.py:84-111) on the host for one clip: int16 (N, 4) FOA audio W,Y,Z,X
    and the event table (E, 4) [frame, class, azi, ele] -> rotated copies.  numpy, a few hundred microseconds per
    clip; used by ``RawAudioDataset`` inside DataLoader workers (the fused device form is ``rot_comb=``).  The int16
    product wraps for -32768 exactly as the reference's numpy multiplication does."""
    import numpy as np
    bits = ROT_BITS[int(comb_no) & 15]
    a = np.array(audio_i16, dtype=np.int16, copy=True)
    if bits & 1: a[:, 1] = (a[:, 1] * np.int16(-1)).astype(np.int16)          # Y
    if bits & 2: a[:, 2] = (a[:, 2] * np.int16(-1)).astype(np.int16)          # Z
    if bits & 4: a[:, 3] = (a[:, 3] * np.int16(-1)).astype(np.int16)          # X
    if bits & 8: a[:, [1, 3]] = a[:, [3, 1]]                                   # X <-> Y
    ev = np.array(events, dtype=np.float64, copy=True).reshape(-1, 4)
    if len(ev):
        pw = -1.0 if comb_no & 2 else 1.0
        dpi = (0.0, 180.0, 90.0, -90.0)[(comb_no >> 2) & 3]
        tw = -1.0 if comb_no & 1 else 1.0
        az = ev[:, 2] * pw + dpi
        az = np.where(az < -180.0, az + 360.0, np.where(az > 180.0, az - 360.0, az))
        ev[:, 2], ev[:, 3] = az, ev[:, 3] * tw
    return a, ev


class SpecAug(object):
    """Same constructor and ``augment(spectrogram)`` as the reference class, plus the batched form."""

    def __init__(self, params: dict, is_valid: bool):
        ac = params["aug_config"]
        self.thresh = ac["spec_augment_thresh"]
        self.time_mask_param = int(ac["spec_augment_time_mask_param"])
        self.freq_mask_param = int(ac["spec_augment_freq_mask_param"])
        self.apply_augment = bool(ac["spec_augment"]) and not is_valid
        self.augment = self._mask if self.apply_augment else self._pass

    # ---- host draws ------------------------------------------------------------------------
    @staticmethod
    def _draw(mask_param: int, axis_len: int):
        """torchaudio.functional.mask_along_axis (p = 1.0): [start, end) on an axis of length axis_len."""
        if mask_param < 1:
            return 0, 0
        value = torch.rand(1) * mask_param
        min_value = torch.rand(1) * (axis_len - value)
        start = int(min_value.long())
        return start, start + int(value.long())

    def draw(self, n_clips: int, n_frames: int, n_mels: int = 64, n_groups: int = 2) -> torch.Tensor:
        """int32 (n_clips, n_groups, 4) = [mel0, mel1, frame0, frame1) in the reference's draw order
        (clip-major, group-minor; gate, width, start per mask).  All zeros when augmentation is off."""
        rects = torch.zeros((n_clips, n_groups, 4), dtype=torch.int32)
        if not self.apply_augment:
            return rects
        for c in range(n_clips):
            for g in range(n_groups):
                if random.random() <= self.thresh:                       # augmentations.py:29-30 -> last axis = mel
                    m0, m1 = self._draw(self.time_mask_param, n_mels)
                    rects[c, g, 0], rects[c, g, 1] = m0, m1
                if random.random() <= self.thresh:                       # augmentations.py:31-32 -> axis -2 = frames
                    f0, f1 = self._draw(self.freq_mask_param, n_frames)
                    rects[c, g, 2], rects[c, g, 3] = f0, f1
        return rects

    # ---- device application ----------------------------------------------------------------
    def apply_rects(self, feat: torch.Tensor, rects: torch.Tensor, groups=FOA_GROUPS) -> torch.Tensor:
        """Zero the drawn intervals of ``feat`` (B, C, T, F) float32 in place; returns ``feat``."""
        require_cuda(feat, "SpecAug")
        if feat.dtype != torch.float32 or feat.dim() != 4 or not feat.is_contiguous():
            raise ValueError("SpecAug expects a contiguous float32 (B, C, T, F) tensor")
        B, Cc, T, F = feat.shape
        G = len(groups)
        if tuple(rects.shape) != (B, G, 4):
            raise ValueError(f"rects must be ({B}, {G}, 4); got {tuple(rects.shape)}")
        for c0, c1 in groups:
            if not 0 <= c0 <= c1 <= Cc:
                raise ValueError("group channel range outside the feature tensor")
        with torch.cuda.device(feat.device):
            d_rects = rects.to(device=feat.device, dtype=torch.int32, non_blocking=True).contiguous()
            d_bounds = torch.tensor(groups, dtype=torch.int32).to(feat.device, non_blocking=True)
            check(_lib.lib().adyolo_spec_mask(ptr(feat), B, Cc, T, F, ptr(d_rects), G, ptr(d_bounds), stream_ptr()),
                  "adyolo_spec_mask")
        return feat

    def augment_batched(self, feat: torch.Tensor, groups=FOA_GROUPS) -> torch.Tensor:
        """Draw + apply for a whole batch (the training-path form: features_batched -> this)."""
        if not self.apply_augment:
            return feat
        return self.apply_rects(feat, self.draw(feat.shape[0], feat.shape[2], feat.shape[3], len(groups)), groups)

    # ---- reference per-clip surface --------------------------------------------------------
    def _pass(self, spectrogram):
        return spectrogram

    def _mask(self, spectrogram: torch.Tensor) -> torch.Tensor:
        """augmentations.py:28-33 for one group (C, T, F) on a CUDA device; returns a masked copy."""
        require_cuda(spectrogram, "SpecAug")
        x = spectrogram.to(torch.float32).contiguous().clone().unsqueeze(0)
        rects = self.draw(1, x.shape[2], x.shape[3], 1)
        return self.apply_rects(x, rects, groups=((0, x.shape[1]),))[0]
