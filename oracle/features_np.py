"""CPU oracle for the SELD feature front end (TEST INFRASTRUCTURE — not a product path).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module.  The product (``ad-yolo_b200``) never does.

What it restates (float64 numpy, same operation order as the reference):

* ``/root/reference/src/utils/utility.py:142-165``  ``audio2stft``
* ``/root/reference/src/utils/utility.py:168-191``  ``stft2melscale``
* ``/root/reference/src/utils/utility.py:194-215``  ``stft2iv``
* ``/root/reference/src/datasets.py:252-292``       ``FeatureLabelProcessor.get_*`` / ``get_feature``
* ``/root/reference/src/preprocess.py:87-130``      ``preprocess_scaler``

The arithmetic the reference delegates to **librosa == 0.8.1** (``requirements.txt:3``; the
package is NOT vendored under /root/reference and is not installable here) is restated from
its published algorithm:

* ``librosa.core.stft(y, n_fft, hop_length, win_length, window)`` with the 0.8.1 defaults
  ``center=True, pad_mode='reflect', dtype=None`` (complex128 for float64 input): reflect-pad
  ``n_fft//2`` both sides, ``1 + len(y)//hop`` frames, periodic window from
  ``scipy.signal.get_window(window, win_length, fftbins=True)``, ``numpy.fft.rfft`` per frame.
* ``librosa.filters.mel(sr, n_fft, n_mels)`` defaults ``fmin=0, fmax=sr/2, htk=False,
  norm='slaney', dtype=float32``.
* ``librosa.power_to_db(S)`` defaults ``ref=1.0, amin=1e-10, top_db=80.0`` (global max).

Parity status: the reference ships no tests/golden vectors for this path and
librosa itself is absent, so the librosa arithmetic is **pinned only by independent
cross-checks** (``torch.stft`` and ``torchaudio.functional.melscale_fbanks``, see
``tests/test_oracle_features.py``); the *structure* (call order, shapes, scaler use) is pinned by
running the unmodified reference ``FeatureLabelProcessor`` on top of this module through the
``librosa`` stub in ``oracle/ref_shims.py`` (``oracle/make_golden.py``).

GCC-PHAT (``gcc_phat``) has no counterpart in the reference at all: it follows the
upstream DCASE baseline ``cls_feature_class.py::_get_gcc`` semantics and is **parity unpinned**.
"""
from __future__ import annotations

import numpy as np

EPS = 1e-8  # utility.py:20 / datasets.py:204


# ----------------------------------------------------------------------------- librosa 0.8.1
def hann_periodic(win_length: int) -> np.ndarray:
    """scipy.signal.get_window('hann', M, fftbins=True) == 0.5 - 0.5 cos(2 pi n / M)."""
    n = np.arange(win_length, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)


def librosa_stft(y: np.ndarray, n_fft: int, hop_length: int, win_length: int,
                 window: str = "hann") -> np.ndarray:
    """librosa.core.stft (0.8.1 defaults) for 1-D float64 ``y`` -> (1+n_fft//2, 1+len//hop) c128."""
    if window not in ("han", "hann", "hanning"):
        raise NotImplementedError("oracle restates the Hann window only (reference yaml: 'han')")
    y = np.asarray(y, dtype=np.float64)
    w = hann_periodic(win_length)
    if win_length < n_fft:  # librosa.util.pad_center
        lpad = (n_fft - win_length) // 2
        w = np.pad(w, (lpad, n_fft - win_length - lpad))
    ypad = np.pad(y, n_fft // 2, mode="reflect")
    n_frames = 1 + (len(ypad) - n_fft) // hop_length
    idx = np.arange(n_fft)[:, None] + hop_length * np.arange(n_frames)[None, :]
    frames = ypad[idx]  # (n_fft, n_frames)
    return np.fft.rfft(w[:, None] * frames, axis=0)


def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def librosa_mel(sr: int, n_fft: int, n_mels: int = 128) -> np.ndarray:
    """librosa.filters.mel (Slaney scale + Slaney area norm) -> (n_mels, 1+n_fft//2) float32."""
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    fftfreqs = np.linspace(0, float(sr) / 2, 1 + n_fft // 2, endpoint=True)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(float(sr) / 2), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]  # float32 *= float64 (computed in f64, stored f32)
    return weights


def librosa_power_to_db(S: np.ndarray, ref: float = 1.0, amin: float = 1e-10,
                        top_db: float | None = 80.0) -> np.ndarray:
    S = np.asarray(S)
    log_spec = 10.0 * np.log10(np.maximum(amin, S))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref))
    if top_db is not None:
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


# ----------------------------------------------------------------------------- reference path
def normalise_int16(audio_i16: np.ndarray) -> np.ndarray:
    """datasets.py:147 / preprocess.py:105."""
    return audio_i16 / 32768.0 + 1e-8


def audio2stft(audio_input, nb_spectra_frames, n_fft, hop_length, win_length, window="han"):
    """utility.py:142-165 -> (T, F, C) complex128."""
    linear_spectra = []
    for ch_idx in range(audio_input.shape[-1]):
        ch_stft = librosa_stft(np.asfortranarray(audio_input[:, ch_idx]), n_fft=n_fft,
                               hop_length=hop_length, win_length=win_length, window=window)
        linear_spectra.append(ch_stft[:, :nb_spectra_frames])
    return np.array(linear_spectra).T


def stft2melscale(linear_spectra, sr, n_fft, mel_bins, top_db=80.0):
    """utility.py:168-191 -> (T, mel, C) float64 log-mel with per-channel global top_db clamp."""
    mel_wts = librosa_mel(sr=sr, n_fft=n_fft, n_mels=mel_bins).T
    mel_spectra = np.zeros((linear_spectra.shape[0], mel_bins, linear_spectra.shape[-1]))
    for ch_idx in range(linear_spectra.shape[-1]):
        magnitude = np.abs(linear_spectra[:, :, ch_idx]) ** 2
        melscale = np.dot(magnitude, mel_wts)
        mel_spectra[:, :, ch_idx] = librosa_power_to_db(melscale, top_db=top_db)
    return mel_spectra


def stft2iv(linear_spectra, sr, n_fft, mel_bins):
    """utility.py:194-215 -> (T, mel, 3) float64 mel-scale FOA intensity vectors."""
    mel_wts = librosa_mel(sr=sr, n_fft=n_fft, n_mels=mel_bins).T
    W = linear_spectra[:, :, 0]
    I = np.real(np.conj(W)[:, :, np.newaxis] * linear_spectra[:, :, 1:])
    E = EPS + (np.abs(W) ** 2 + ((np.abs(linear_spectra[:, :, 1:]) ** 2).sum(-1)) / 3.0)
    I_norm = I / E[:, :, np.newaxis]
    I_norm_mel = np.transpose(np.dot(np.transpose(I_norm, (0, 2, 1)), mel_wts), (0, 2, 1))
    if np.isnan(I_norm_mel).any():
        raise FloatingPointError("Feature extraction is generating nan outputs")
    return I_norm_mel


def gcc_phat(linear_spectra, n_fft, nb_lags):
    """Upstream seld-dcase2022 ``_get_gcc`` semantics (NOT in the reference; parity unpinned).

    For each microphone pair m<n: R = conj(X_m) X_n, cc = irfft(exp(1j*angle(R))) of length
    n_fft, output lags [-nb_lags/2, nb_lags/2): concat(cc[-nb_lags//2:], cc[:nb_lags//2]).
    -> (T, nb_lags, C*(C-1)/2)
    """
    T, F, C = linear_spectra.shape
    out = []
    for m in range(C):
        for n in range(m + 1, C):
            R = np.conj(linear_spectra[:, :, m]) * linear_spectra[:, :, n]
            # upstream's np.angle(R) of an exactly vanishing R is 0 or pi depending on the SIGN of the
            # zero (conj(0) * x can be -0.0): an accident, pinned here to phase 0 (what the kernel does)
            ph = np.where(R == 0, 1.0 + 0.0j, np.exp(1.0j * np.angle(R)))
            cc = np.fft.irfft(ph, n=n_fft, axis=1)
            out.append(np.concatenate((cc[:, -nb_lags // 2:], cc[:, :nb_lags // 2]), axis=-1))
    return np.stack(out, axis=-1)


def features_foa(audio: np.ndarray, sr=24000, n_fft=1200, hop_length=600, win_length=1200,
                 mel_bins=64, window="han", scaler=None, label_hop_len=2400):
    """datasets.py:281-292 ``get_feature`` on normalised float64 ``audio`` (N, 4).

    Returns ([MEL (T,64,4), IV (T,64,3)], nb_label_frames); standardised when ``scaler`` given.
    """
    nb_feature_frames = int(len(audio) / float(hop_length))
    nb_label_frames = int(len(audio) / float(label_hop_len))
    spec = audio2stft(audio, nb_feature_frames, n_fft, hop_length, win_length, window)
    MEL = stft2melscale(spec, sr, n_fft, mel_bins)
    IV = stft2iv(spec, sr, n_fft, mel_bins)
    if scaler is not None:
        MEL = (MEL - scaler["MEL"]["mean"]) / scaler["MEL"]["std"]
        IV = (IV - scaler["IV"]["mean"]) / scaler["IV"]["std"]
    return [MEL, IV], nb_label_frames


def features_foa_stack(audio_i16: np.ndarray, scaler=None, **kw) -> np.ndarray:
    """int16 (N,4) clip -> (7, T, 64) float64 in the layout Dataset.__getitem__ builds
    (datasets.py:147-160: normalise, get_feature, permute(2,0,1), cat) minus augmentation."""
    (MEL, IV), _ = features_foa(normalise_int16(audio_i16), scaler=scaler, **kw)
    return np.concatenate([MEL.transpose(2, 0, 1), IV.transpose(2, 0, 1)], axis=0)


def features_mic_stack(audio_i16: np.ndarray, scaler=None, sr=24000, n_fft=1200, hop_length=600,
                       win_length=1200, mel_bins=64, window="han") -> np.ndarray:
    """MIC format: 4 log-mel + 6 GCC-PHAT -> (10, T, 64) float64 (parity unpinned, see header)."""
    audio = normalise_int16(audio_i16)
    T = int(len(audio) / float(hop_length))
    spec = audio2stft(audio, T, n_fft, hop_length, win_length, window)
    MEL = stft2melscale(spec, sr, n_fft, mel_bins)
    GCC = gcc_phat(spec, n_fft, mel_bins)
    if scaler is not None:
        MEL = (MEL - scaler["MEL"]["mean"]) / scaler["MEL"]["std"]
        GCC = (GCC - scaler["GCC"]["mean"]) / scaler["GCC"]["std"]
    return np.concatenate([MEL.transpose(2, 0, 1), GCC.transpose(2, 0, 1)], axis=0)


def scaler_stats(clips_i16, fmt="foa", **kw):
    """preprocess.py:103-127: per-file features (clamp per file), concatenate, mean/std/max/min
    over frames -> dict of (1, 64, C) float64 arrays."""
    A, Bk = [], []
    for clip in clips_i16:
        if fmt == "foa":
            (MEL, SEC), _ = features_foa(normalise_int16(clip), **kw)
        else:
            audio = normalise_int16(clip)
            T = int(len(audio) / 600.0)
            spec = audio2stft(audio, T, 1200, 600, 1200, "han")
            MEL = stft2melscale(spec, 24000, 1200, 64)
            SEC = gcc_phat(spec, 1200, 64)
        A.append(MEL)
        Bk.append(SEC)
    A = np.concatenate(A, axis=0)
    Bk = np.concatenate(Bk, axis=0)
    key = "IV" if fmt == "foa" else "GCC"
    out = {"MEL": {}, key: {}}
    for name, stack in (("MEL", A), (key, Bk)):
        out[name]["mean"] = stack.mean(0, keepdims=True)
        out[name]["std"] = stack.std(0, keepdims=True)
        out[name]["max"] = stack.max(0, keepdims=True)
        out[name]["min"] = stack.min(0, keepdims=True)
    return out
