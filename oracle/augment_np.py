"""CPU oracle for the FOA rotation augmentation (TEST INFRASTRUCTURE — not a product path).

Restates ``/root/reference/src/utils/augmentations.py:36-111`` (``RotationAug``): the table of the
16 channel sign / swap combinations with their label transforms, ``_rotate`` on int16 audio
(numpy int16 arithmetic, so ``-1 * -32768`` wraps exactly like the reference) and on the label
dict.  Pinned against the unmodified reference class by ``oracle/make_golden.py`` ->
``tests/golden/rotation.npz``.
"""
from __future__ import annotations

import copy

import numpy as np

# (yzx_weight, xy_swap, pi_weight, d_pi, theta_weight)   augmentations.py:46-70
ROTATION_COMBINATION = [
    ([1, 1, 1], False, 1, 0, 1), ([1, -1, 1], False, 1, 0, -1),
    ([-1, 1, 1], False, -1, 0, 1), ([-1, -1, 1], False, -1, 0, -1),
    ([-1, 1, -1], False, 1, 180, 1), ([-1, -1, -1], False, 1, 180, -1),
    ([1, 1, -1], False, -1, 180, 1), ([1, -1, -1], False, -1, 180, -1),
    ([-1, 1, 1], True, 1, 90, 1), ([-1, -1, 1], True, 1, 90, -1),
    ([1, 1, 1], True, -1, 90, 1), ([1, -1, 1], True, -1, 90, -1),
    ([1, 1, -1], True, 1, -90, 1), ([1, -1, -1], True, 1, -90, -1),
    ([-1, 1, -1], True, -1, -90, 1), ([-1, -1, -1], True, -1, -90, -1),
]


def rotate(audio: np.ndarray, label: dict, comb_no: int):
    """augmentations.py:84-111 (does not modify its inputs)."""
    w, swap, pw, dpi, tw = ROTATION_COMBINATION[int(comb_no)]
    audio = audio.copy()
    for ch in range(1, 4):
        audio[:, ch] = audio[:, ch] * w[ch - 1]
    if swap:
        audio = audio[:, [0, 3, 2, 1]]
    label = copy.deepcopy(label)
    for frame_idx in label.keys():
        for ev in label[frame_idx]:
            pi, theta = ev[-2], ev[-1]
            pi = pi * pw + dpi
            theta = theta * tw
            if pi < -180:
                pi = pi + 360
            elif pi > 180:
                pi = pi - 360
            ev[-2], ev[-1] = pi, theta
    return audio, label


def rotate_events(events: np.ndarray, comb_of_batch: np.ndarray) -> np.ndarray:
    """Array form: events (E,5) [batch, frame, class, azi, ele] with a combination per clip."""
    ev = np.array(events, dtype=np.float64, copy=True)
    for i in range(len(ev)):
        _, _, pw, dpi, tw = ROTATION_COMBINATION[int(comb_of_batch[int(ev[i, 0])])]
        pi = ev[i, 3] * pw + dpi
        if pi < -180:
            pi += 360
        elif pi > 180:
            pi -= 360
        ev[i, 3], ev[i, 4] = pi, ev[i, 4] * tw
    return ev


# ------------------------------------------------------------------------------------------------
def specaug_rects(n_clips: int, n_frames: int, n_mels: int, thresh: float, time_mask_param: int,
                  freq_mask_param: int, n_groups: int = 2):
    """Restates ``/root/reference/src/utils/augmentations.py:6-33`` (``SpecAug._mask``) as applied by
    ``datasets.py:158-160`` to each feature group permuted to (C, T, F), together with the arithmetic
    of torchaudio 2.x ``functional.mask_along_axis`` (library code the reference calls; p = 1.0,
    mask value 0):  ``TimeMasking`` acts on the LAST axis, which for (C, T, F) is the mel axis, with
    ``spec_augment_time_mask_param``; ``FrequencyMasking`` acts on axis -2 = the frame axis with
    ``spec_augment_freq_mask_param``.  Draw order per group: ``random.random()`` gate, then
    ``torch.rand(1)`` for the width and ``torch.rand(1)`` for the start (float32 arithmetic, ``.long()``
    truncation).  Consumes the global python / torch CPU RNGs exactly like the reference.

    -> int32 (n_clips, n_groups, 4) = [mel0, mel1, frame0, frame1] (empty interval = no mask).
    Pinned against the unmodified reference class by ``tests/golden/specaug.npz``."""
    import random

    import torch

    def draw(mask_param, axis_len):
        if mask_param < 1:
            return 0, 0
        value = torch.rand(1) * mask_param
        min_value = torch.rand(1) * (axis_len - value)
        start = int(min_value.long())
        return start, start + int(value.long())

    out = np.zeros((n_clips, n_groups, 4), np.int32)
    for c in range(n_clips):
        for g in range(n_groups):
            if random.random() <= thresh:
                out[c, g, 0:2] = draw(time_mask_param, n_mels)
            if random.random() <= thresh:
                out[c, g, 2:4] = draw(freq_mask_param, n_frames)
    return out


def specaug_apply(feat: np.ndarray, rects: np.ndarray, groups=((0, 4), (4, 7))) -> np.ndarray:
    """feat (B, C, T, F): zero the drawn intervals of every channel of each group."""
    out = feat.copy()
    for b in range(feat.shape[0]):
        for g, (c0, c1) in enumerate(groups):
            m0, m1, f0, f1 = (int(v) for v in rects[b, g])
            out[b, c0:c1, :, m0:m1] = 0
            out[b, c0:c1, f0:f1, :] = 0
    return out
