"""Pin the rotation-augmentation oracle against the unmodified reference RotationAug."""
import numpy as np
import pytest

from oracle import assign_np, augment_np


@pytest.fixture(scope="module")
def A():
    import __graft_entry__ as g
    g.build()
    import adyolo_b200
    return adyolo_b200


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


def _rects_to_masks(rects, T=200, F=64):
    m = np.zeros((rects.shape[0], rects.shape[1], T, F), np.uint8)
    for c in range(rects.shape[0]):
        for g in range(rects.shape[1]):
            m0, m1, f0, f1 = (int(v) for v in rects[c, g])
            m[c, g, :, m0:m1] = 1
            m[c, g, f0:f1, :] = 1
    return m


def test_specaug_oracle_and_host_draws_match_reference(gold, A):
    """The unmodified reference SpecAug (torchaudio Time/FrequencyMasking on (C,T,F)) was run on
    tensors of ones with python/torch seeds 11 -> tests/golden/specaug.npz.  Both the oracle
    restatement and the product's host-side draw must reproduce its masks from the same seeds
    (host logic only: no CUDA needed)."""
    import random
    import torch
    g = gold("specaug.npz")
    seed, n = int(g["seed"]), g["masked"].shape[0]
    random.seed(seed); torch.manual_seed(seed)
    r_oracle = augment_np.specaug_rects(n, 200, 64, float(g["thresh"]), int(g["time_mask_param"]), int(g["freq_mask_param"]))
    np.testing.assert_array_equal(_rects_to_masks(r_oracle), g["masked"])
    params = {"aug_config": {"spec_augment": True, "spec_augment_thresh": float(g["thresh"]),
                             "spec_augment_time_mask_param": int(g["time_mask_param"]),
                             "spec_augment_freq_mask_param": int(g["freq_mask_param"])}}
    random.seed(seed); torch.manual_seed(seed)
    r_prod = A.SpecAug(params, is_valid=False).draw(n, 200, 64).numpy()
    np.testing.assert_array_equal(r_prod, r_oracle)
    assert g["masked"].any() and (g["masked"].reshape(n, 2, -1).max(-1) == 0).any()     # some groups masked, some not
    # validation / disabled: nothing drawn, RNG untouched
    random.seed(seed); torch.manual_seed(seed)
    assert not A.SpecAug(params, is_valid=True).draw(n, 200, 64).any()
    assert random.random() == random.Random(seed).random()


def test_rotation_draw_sequence_matches_reference(gold, A):
    """RotationAug.draw consumes python's RNG exactly like the reference class (golden: the
    combinations the unmodified class picked under random.seed(21))."""
    import random
    g = gold("rotation.npz")
    params = {"aug_config": {"rotation_augment": True}}
    random.seed(int(g["draw_seed"]))
    got = A.RotationAug(params, is_valid=False).draw(len(g["draws"]), "cpu")
    np.testing.assert_array_equal(got.numpy().astype(np.int64), g["draws"])
    assert len(set(g["draws"].tolist())) > 8
    assert A.RotationAug(params, is_valid=True).draw(4, "cpu") is None
