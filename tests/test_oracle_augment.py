"""Pin the rotation-augmentation oracle against the unmodified reference RotationAug."""
import numpy as np

from oracle import assign_np, augment_np


def _label(frames, events):
    lab = {}
    for f, e in zip(frames, events):
        lab.setdefault(int(f), []).append([int(e[0]), int(e[1]), float(e[2]), float(e[3])])
    return lab


def test_rotation_oracle_matches_reference(gold):
    g = gold("rotation.npz")
    lab = _label(g["label_frames"], g["label_events"])
    for c in range(16):
        a2, l2 = augment_np.rotate(g["snippet"], lab, c)
        np.testing.assert_array_equal(a2, g[f"audio_{c}"])
        ev = np.asarray([e for fr, evs in l2.items() for e in evs], np.float64)
        np.testing.assert_array_equal(ev, g[f"events_{c}"])
        rows = assign_np.get_yolo_label(l2, 20)
        np.testing.assert_array_equal(np.asarray(rows, np.float64).reshape(-1, 6), g[f"rows_{c}"].reshape(-1, 6))
    assert g["audio_2"][0, 1] == -32768          # the reference's int16 wrap of -1 * -32768
