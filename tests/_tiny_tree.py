"""A tiny synthetic DCASE-style tree (foa_dev / metadata_dev, the reference's directory layout,
datasets.py:36-56) for the drop-in tests: int16 4-channel wavs + polar csv labels + a scaler pickle."""
import os
import pickle

import numpy as np


def params_for(root, nb_classes=12, device="cpu"):
    from oracle.loss_torch import default_params
    p = default_params(nb_classes, device)
    p["data_config"].update({"data_pth": str(root) + os.sep, "chunk_window_s": 2, "chunk_stride_s": 1})
    p["train_config"].update({"batch_size": 2, "nb_iters": 5, "conf_thresh": 0.3, "clss_thresh": 0.3, "unify_thresh": 15.0,
                              "nms": "conn-merge", "lr": 1e-3, "weight_decay": 0.0, "optim": "Adam"})
    p["aug_config"] = {"rotation_augment": False, "spec_augment": False, "spec_augment_thresh": 0.5,
                       "spec_augment_time_mask_param": 8, "spec_augment_freq_mask_param": 8}
    p["args"]["quick_test"] = True
    return p


def build(root, scaler, n_train=12, n_test=2, seed=0):
    """train chunks of 2 s (20 label frames), test clips of 3 s; every clip has events."""
    import scipy.io.wavfile as wav
    rng = np.random.default_rng(seed)
    sets = {"dev-train-chunked_2s_1s": (n_train, 2.0), "dev-test": (n_test, 3.0)}
    for sub, (n, secs) in sets.items():
        wd, cd = os.path.join(root, "foa_dev", sub), os.path.join(root, "metadata_dev", sub)
        os.makedirs(wd, exist_ok=True)
        os.makedirs(cd, exist_ok=True)
        for i in range(n):
            name = f"fold1_room1_mix{i + 1:03d}" + ("_chunk001" if "chunk" in sub else "")
            N = int(24000 * secs)
            x = rng.standard_normal((N, 4)) * 800.0
            t = np.arange(N) / 24000.0
            x += 3000.0 * np.sin(2 * np.pi * rng.uniform(200, 4000) * t)[:, None] * rng.uniform(-1, 1, 4)[None, :]
            wav.write(os.path.join(wd, name + ".wav"), 24000, np.clip(np.round(x), -32767, 32767).astype(np.int16))
            with open(os.path.join(cd, name + ".csv"), "w") as f:
                for fr in range(int(secs * 10)):
                    for src in range(int(rng.integers(0, 3))):
                        f.write("{},{},{},{},{}\n".format(fr, int(rng.integers(0, 12)), src, int(rng.integers(-180, 181)),
                                                          int(rng.integers(-80, 81))))
                f.write("{},{},{},{},{}\n".format(0, 3, 0, 45, 10))      # at least one event per clip
    with open(os.path.join(root, "scaler_wts.pkl"), "wb") as f:
        pickle.dump(scaler, f)
