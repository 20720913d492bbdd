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
