"""Import the UNMODIFIED reference modules from /root/reference/src in this container
(TEST INFRASTRUCTURE; used only by oracle/make_golden.py and the CPU-only pinning tests that
skip when /root/reference is absent, e.g. on the GPU box).

The reference cannot be imported as-is here because third-party packages are missing
(librosa 0.8.1, ruamel.yaml, neptune, soundfile) and ``utils/seld_metrics.py:4`` uses the
removed ``np.float``.  The shims below supply *module objects only*; no reference source is
modified or copied.  ``librosa`` is replaced by the restated algorithms in
``oracle/features_np.py`` so that the reference's own ``FeatureLabelProcessor`` /
``audio2stft`` / ``stft2melscale`` / ``stft2iv`` code runs on top of them.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REF_SRC = "/root/reference/src"


def reference_available() -> bool:
    return os.path.isdir(REF_SRC)


def install():
    """Make ``import datasets``, ``import models.loss``, ``import utils.utility`` work."""
    if not reference_available():
        raise RuntimeError("/root/reference is not present (GPU box?)")
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here))
    from oracle import features_np as F

    if not hasattr(np, "float"):
        np.float = float  # seld_metrics.py:4

    librosa = types.ModuleType("librosa")
    librosa.core = types.ModuleType("librosa.core")
    librosa.filters = types.ModuleType("librosa.filters")

    def _stft(y, n_fft=2048, hop_length=None, win_length=None, window="hann", **kw):
        return F.librosa_stft(y, n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=window)

    librosa.core.stft = _stft
    librosa.stft = _stft
    librosa.filters.mel = lambda sr, n_fft, n_mels=128, **kw: F.librosa_mel(sr, n_fft, n_mels)
    librosa.power_to_db = F.librosa_power_to_db
    sys.modules.setdefault("librosa", librosa)
    sys.modules.setdefault("librosa.core", librosa.core)
    sys.modules.setdefault("librosa.filters", librosa.filters)

    for name in ("ruamel", "ruamel.yaml", "neptune", "neptune.new", "soundfile", "mat73"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["ruamel"], "yaml"):
        sys.modules["ruamel"].yaml = sys.modules["ruamel.yaml"]
    if not hasattr(sys.modules["neptune"], "new"):
        sys.modules["neptune"].new = sys.modules["neptune.new"]

    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)


def ref_params(nb_classes=12, device="cpu", data_pth="/root/reference/data/DCASE2021_SELD/"):
    from oracle.loss_torch import default_params
    p = default_params(nb_classes, device)
    p["data_config"]["data_pth"] = data_pth
    p["aug_config"] = {"rotation_augment": False, "spec_augment": False, "spec_augment_thresh": 0.5,
                       "spec_augment_time_mask_param": 8, "spec_augment_freq_mask_param": 8}
    return p
