"""Feature front end: Python mirror of the reference call surface over the CUDA C ABI.

Reference surface mirrored (same names, argument meaning, return shapes):
  * ``utils/utility.py:142-215``  audio2stft / stft2melscale / stft2iv
  * ``datasets.py:187-292``       FeatureLabelProcessor (feature half)
and the batched entry the training path needs (DataLoader workers cannot use CUDA,
so workers hand int16 audio to the main process): ``features_batched``.

No CPU fallback: every function raises RuntimeError without the built library + a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
import pickle

import numpy as np
import torch

from . import _lib
from ._lib import FrontendCfg, check, ptr, require_cuda, stream_ptr

_SCALER_KEYS = ("MEL", "IV")
FRONTEND_KERNEL = "fe2_foa_kernel"          # name of the fused front-end kernel (bench.py's roofline / ncu filters)


def _cfg(sr=24000, n_fft=1200, hop_length=600, win_length=1200, mel_bins=64, n_channels=4,
         dc_offset=1e-8, top_db=80.0) -> FrontendCfg:
    return FrontendCfg(int(sr), int(n_fft), int(hop_length), int(win_length), int(mel_bins), int(n_channels),
                       float(dc_offset), float(top_db))


def _check_window(window):
    if window not in ("han", "hann", "hanning"):
        raise NotImplementedError(f"window {window!r}: only the reference's Hann window is implemented")


def mel_filterbank(sr: int, n_fft: int, n_mels: int) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels) (0.8.1 defaults) computed by the library's host code."""
    out = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    check(_lib.lib().adyolo_mel_filterbank(sr, n_fft, n_mels, out.ctypes.data_as(C.c_void_p)), "adyolo_mel_filterbank")
    return out


def _scaler_to_device(scaler, keys, device):
    """{'MEL': {'mean','std'}, 'IV'|'GCC': {...}} with (1,64,C) arrays -> (mean, inv_std) (Ctot,64) f32."""
    means, stds = [], []
    for k in keys:
        means.append(np.asarray(scaler[k]["mean"], np.float64)[0].T)   # (C,64)
        stds.append(np.asarray(scaler[k]["std"], np.float64)[0].T)
    mean = torch.from_numpy(np.concatenate(means, 0).astype(np.float32)).to(device).contiguous()
    istd = torch.from_numpy((1.0 / np.concatenate(stds, 0)).astype(np.float32)).to(device).contiguous()
    return mean, istd


_ws_cache = {}


def _workspace(nbytes: int, device) -> torch.Tensor:
    key = (str(device), torch.cuda.current_stream(device).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def features_batched(audio_i16: torch.Tensor, scaler_dev=None, out: torch.Tensor | None = None,
                     apply_topdb: bool = True, cfg: FrontendCfg | None = None,
                     check_nan: bool = False, timing_events: list | None = None,
                     rot_comb: torch.Tensor | None = None) -> torch.Tensor:
    """int16 PCM (B, N, 4) on a CUDA device -> features (B, 7, T, 64) float32 on the same device.

    Equivalent to, per clip: ``audio/32768.0 + 1e-8`` -> ``FeatureLabelProcessor.get_feature`` ->
    ``permute(2,0,1)`` + ``cat`` (datasets.py:147-160) without augmentation.  ``scaler_dev`` is the
    ``(mean, inv_std)`` pair from ``FeatureLabelProcessor.scaler_device`` or None (raw dB / IV).
    ``timing_events``: when a list is given, a (start, end) pair of CUDA events bracketing the fused
    front-end kernel alone is appended (bench.py's roofline measurement); the top_db pass is then
    issued through ``adyolo_features_foa_clamp``.
    ``rot_comb``: optional int8 tensor (B,) on the device with the ``RotationAug`` combination number
    (0..15, ``utils/augmentations.py:46-70``) of each clip: the augmentation is fused into the kernel.
    """
    require_cuda(audio_i16, "features_batched")
    if audio_i16.dtype != torch.int16 or audio_i16.dim() != 3 or audio_i16.shape[-1] != 4:
        raise ValueError("features_batched expects an int16 tensor of shape (B, N, 4)")
    audio_i16 = audio_i16.contiguous()
    B, N, _ = audio_i16.shape
    T = N // 600
    cfg = cfg or _cfg()
    L = _lib.lib()
    with torch.cuda.device(audio_i16.device):
        if out is None:
            out = torch.empty((B, 7, T, 64), dtype=torch.float32, device=audio_i16.device)
        elif out.shape != (B, 7, T, 64) or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError("out must be a contiguous float32 (B, 7, T, 64) tensor")
        nbytes = L.adyolo_frontend_workspace_bytes(C.byref(cfg), B, N)
        ws = _workspace(nbytes, audio_i16.device)
        mean, istd = scaler_dev if scaler_dev is not None else (None, None)
        if rot_comb is not None:
            if rot_comb.dtype != torch.int8 or rot_comb.shape != (B,) or not rot_comb.is_cuda:
                raise ValueError("rot_comb must be an int8 CUDA tensor of shape (B,)")
            rot_comb = rot_comb.contiguous()
        if timing_events is None:
            check(L.adyolo_features_foa_rot(ptr(audio_i16), B, N, C.byref(cfg), ptr(mean), ptr(istd), ptr(rot_comb),
                                            ptr(out), ptr(ws), 1 if apply_topdb else 0, stream_ptr()),
                  "adyolo_features_foa")
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            check(L.adyolo_features_foa_rot(ptr(audio_i16), B, N, C.byref(cfg), ptr(mean), ptr(istd), ptr(rot_comb),
                                            ptr(out), ptr(ws), 0, stream_ptr()), "adyolo_features_foa")
            e1.record()
            timing_events.append((e0, e1))
            if apply_topdb:
                check(L.adyolo_features_foa_clamp(ptr(out), B, N, C.byref(cfg), ptr(mean), ptr(istd), ptr(ws),
                                                  stream_ptr()), "adyolo_features_foa_clamp")
        if check_nan and int(ws[:4].view(torch.int32).item()) & 1:
            raise FloatingPointError("Feature extraction is generating nan outputs")  # datasets.py:277
    return out


# ------------------------------------------------------------------------------------------------
# utility.py function surface (numpy in / numpy out, single clip)
def _audio_to_device(audio_input: np.ndarray):
    """Reference callers pass the normalised float64 clip (x/32768 + 1e-8).  When it is exactly an
    int16 clip in that form the int16 kernels are used (bit-identical input to the reference's);
    otherwise the float32-input kernels."""
    a = np.asarray(audio_input)
    if a.ndim != 2 or a.shape[1] != 4:
        raise ValueError("audio must have shape (N, 4)")
    if a.dtype == np.int16:
        return torch.from_numpy(np.ascontiguousarray(a)).cuda(), 0
    h = np.rint((a.astype(np.float64) - 1e-8) * 32768.0)
    if np.all(np.abs(h) <= 32768) and np.array_equal(h / 32768.0 + 1e-8, a) and h.min() >= -32768 and h.max() <= 32767:
        return torch.from_numpy(h.astype(np.int16)).cuda(), 0
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda(), 1


def _stft_dev(audio_dev: torch.Tensor, dtype_code: int, nb_frames: int, cfg: FrontendCfg) -> torch.Tensor:
    N = audio_dev.shape[-2]
    T = N // 600
    x = audio_dev.reshape(-1, N, 4)
    B = x.shape[0]
    spec = torch.empty((B, T, 601, 4), dtype=torch.complex64, device=x.device)
    check(_lib.lib().adyolo_stft(ptr(x), dtype_code, B, N, C.byref(cfg), ptr(spec), stream_ptr()), "adyolo_stft")
    return spec[:, :nb_frames]


def audio2stft(audio_input: np.ndarray, nb_spectra_frames: int, n_fft: int, hop_length: int, win_length: int,
               window: str = "han") -> np.ndarray:
    """utility.py:142-165.  (N, 4) audio -> (T, F, C) complex spectrogram (complex128 array whose
    values were computed in FP32 on the GPU)."""
    require_cuda(None, "audio2stft")
    _check_window(window)
    cfg = _cfg(n_fft=n_fft, hop_length=hop_length, win_length=win_length)
    dev, code = _audio_to_device(audio_input)
    spec = _stft_dev(dev, code, nb_spectra_frames, cfg)[0]
    return spec.cpu().numpy().astype(np.complex128)


def _spec_to_device(linear_spectra: np.ndarray) -> torch.Tensor:
    s = np.ascontiguousarray(np.asarray(linear_spectra), dtype=np.complex64)
    if s.ndim != 3 or s.shape[1] != 601:
        raise NotImplementedError("spectrogram must be (T, 601, C): only n_fft=1200 is compiled in")
    return torch.from_numpy(s).cuda()


def _strides_TFC(T, Cc):
    # output (T, 64, C): element (b,c,t,j) at t*64*C + j*C + c
    return (C.c_int64 * 4)(T * 64 * Cc, 1, 64 * Cc, Cc)


def stft2melscale(linear_spectra: np.ndarray, sr: int, n_fft: int, mel_bins: int) -> np.ndarray:
    """utility.py:168-191 -> (T, mel_bins, C) float64 log-mel (top_db clamp per channel)."""
    require_cuda(None, "stft2melscale")
    cfg = _cfg(sr=sr, n_fft=n_fft, hop_length=n_fft // 2, win_length=n_fft, mel_bins=mel_bins)
    spec = _spec_to_device(linear_spectra)
    T, _, Cs = spec.shape
    if Cs > 4:
        raise NotImplementedError("at most 4 channels")
    out = torch.empty((T, 64, Cs), dtype=torch.float32, device=spec.device)
    gws = torch.empty(max(Cs, 1), dtype=torch.int32, device=spec.device)
    check(_lib.lib().adyolo_logmel_from_stft(ptr(spec), 1, T, Cs, Cs, C.byref(cfg), None, None, ptr(out),
                                             _strides_TFC(T, Cs), ptr(gws), 1, stream_ptr()), "adyolo_logmel_from_stft")
    return out.cpu().numpy().astype(np.float64)


def stft2iv(linear_spectra: np.ndarray, sr: int, n_fft: int, mel_bins: int) -> np.ndarray:
    """utility.py:194-215 -> (T, mel_bins, 3) float64 mel-scale FOA intensity vectors."""
    require_cuda(None, "stft2iv")
    cfg = _cfg(sr=sr, n_fft=n_fft, hop_length=n_fft // 2, win_length=n_fft, mel_bins=mel_bins)
    spec = _spec_to_device(linear_spectra)
    T, _, Cs = spec.shape
    if Cs != 4:
        raise ValueError("FOA intensity vectors need the 4 channels W,Y,Z,X")
    out = torch.empty((T, 64, 3), dtype=torch.float32, device=spec.device)
    flags = torch.zeros(1, dtype=torch.int32, device=spec.device)
    check(_lib.lib().adyolo_iv_from_stft(ptr(spec), 1, T, C.byref(cfg), None, None, ptr(out), _strides_TFC(T, 3),
                                         ptr(flags), stream_ptr()), "adyolo_iv_from_stft")
    if int(flags.item()) & 1:
        raise FloatingPointError("Feature extraction is generating nan outputs")  # utility.py:213
    return out.cpu().numpy().astype(np.float64)


def features_mic_batched(audio_i16: torch.Tensor, scaler_dev=None, apply_topdb: bool = True) -> torch.Tensor:
    """MIC format (absent from the reference, upstream DCASE-baseline semantics):
    int16 (B, N, 4) -> (B, 10, T, 64) = 4 log-mel + 6 GCC-PHAT (64 lags)."""
    require_cuda(audio_i16, "features_mic_batched")
    if audio_i16.dtype != torch.int16 or audio_i16.dim() != 3 or audio_i16.shape[-1] != 4:
        raise ValueError("features_mic_batched expects an int16 tensor of shape (B, N, 4)")
    audio_i16 = audio_i16.contiguous()
    B, N, _ = audio_i16.shape
    T = N // 600
    cfg = _cfg()
    L = _lib.lib()
    with torch.cuda.device(audio_i16.device):
        spec = torch.empty(L.adyolo_mic_spec_bytes(B, N), dtype=torch.uint8, device=audio_i16.device)
        out = torch.empty((B, 10, T, 64), dtype=torch.float32, device=audio_i16.device)
        mean, istd = scaler_dev if scaler_dev is not None else (None, None)
        ws = _workspace(L.adyolo_frontend_workspace_bytes(C.byref(cfg), B, N), audio_i16.device)
        # fused front end (log-mel of the 4 channels + channel spectra) -> tcgen05 GCC-PHAT lag transform
        check(L.adyolo_features_mic_gcc(ptr(audio_i16), B, N, C.byref(cfg), ptr(mean), ptr(istd), ptr(out), ptr(spec),
                                        ptr(ws), 1 if apply_topdb else 0, stream_ptr()), "adyolo_features_mic_gcc")
    return out


# ------------------------------------------------------------------------------------------------
class FeatureLabelProcessor:
    """datasets.py:187-292 (feature half) + :457-482 (adyolo label half, see labels.py).

    ``params`` is the reference's dict-of-dicts; the keys read are the ones the reference reads
    (SURVEY §8(b)).  ``scaler`` may be passed directly instead of ``<data_pth>/scaler_wts.pkl``.
    """

    def __init__(self, params: dict, scaler: dict | None = None, device=None):
        dc = params["data_config"]
        self.nb_classes = dc["nb_classes"]
        self.sr = dc["sr"]
        self.hop_length = dc["hop_length"]
        self.win_length = dc["win_length"]
        self.n_fft = dc["n_fft"]
        self.window = dc["window"]
        self.mel_bins = dc["mel_bins"]
        self.label_hop_len_s = dc["label_hop_len_s"]
        self.label_hop_len = int(dc["sr"] * dc["label_hop_len_s"])
        self.eps = 1e-8
        _check_window(self.window)
        require_cuda(None, "FeatureLabelProcessor")
        self.device = torch.device(device if device is not None else params.get("args", {}).get("device", "cuda"))
        if self.device.type != "cuda":
            self.device = torch.device("cuda")
        self._cfg = _cfg(sr=self.sr, n_fft=self.n_fft, hop_length=self.hop_length, win_length=self.win_length,
                         mel_bins=self.mel_bins)
        self.mel_wts = mel_filterbank(self.sr, self.n_fft, self.mel_bins).T   # datasets.py:203
        if scaler is None:
            with open(os.path.join(dc["data_pth"], "scaler_wts.pkl"), "rb") as f:   # datasets.py:206
                scaler = pickle.load(f)
        self.scaler = scaler
        self.scaler_device = _scaler_to_device(scaler, _SCALER_KEYS, self.device)

        loss = params["args"]["loss"]
        if loss == "adyolo":
            from .labels import GridSpec
            tc = params["train_config"]
            self.grid = GridSpec(self.nb_classes, tc["nb_anchors"], tc["grid_size"], tc["g_overlap"],
                                 tc.get("train_unify", [45., 25., 10.]), tc.get("loss_gains"))
            self.grid_size = np.array(tc["grid_size"])
            self.nb_grids = self.grid.nb_grids
            self.g_overlap = tc["g_overlap"]
            self.nb_anchors = tc["nb_anchors"]
            self.get_label = self.get_yolo_label
        elif loss in ("seddoa", "masked-seddoa", "accdoa", "adpit"):
            raise NotImplementedError(f"loss: {loss} (only --loss adyolo is on the B200 hot path)")
        else:
            raise NotImplementedError("loss: {}".format(loss))   # datasets.py:241

    # ---- feature half
    def get_feature_label(self, audio, label):
        feature_stack, nb_label_frames = self.get_feature(audio)
        return feature_stack, self.get_label(label, nb_label_frames)

    def get_stft_spectrogram(self, audio_input, nb_feature_frames):
        return audio2stft(audio_input, nb_feature_frames, self.n_fft, self.hop_length, self.win_length, self.window)

    def get_logmel_spectrogram(self, linear_spectra):
        return stft2melscale(linear_spectra, self.sr, self.n_fft, self.mel_bins)

    def get_melscale_foa_intensity_vectors(self, linear_spectra):
        return stft2iv(linear_spectra, self.sr, self.n_fft, self.mel_bins)

    def get_feature(self, audio):
        """datasets.py:281-292: normalised (N,4) clip -> ([MEL (T,64,4), IV (T,64,3)], nb_label_frames)."""
        nb_feature_frames = int(len(audio) / float(self.hop_length))
        nb_label_frames = int(len(audio) / float(self.label_hop_len))
        dev, code = _audio_to_device(audio)
        with torch.cuda.device(self.device):
            dev = dev.to(self.device)
            if code == 0:
                feat = features_batched(dev[None], self.scaler_device, cfg=self._cfg, check_nan=True)[0]
            else:
                feat = self._features_from_float(dev)
        f = feat.cpu().numpy().astype(np.float64)                         # (7, T, 64)
        MEL = np.ascontiguousarray(f[:4].transpose(1, 2, 0))[:nb_feature_frames]
        IV = np.ascontiguousarray(f[4:].transpose(1, 2, 0))[:nb_feature_frames]
        return [MEL, IV], nb_label_frames

    def _features_from_float(self, audio_f32: torch.Tensor) -> torch.Tensor:
        """float32-input route: STFT kernel -> log-mel / IV kernels, written as (7, T, 64)."""
        N = audio_f32.shape[0]
        T = N // 600
        L = _lib.lib()
        spec = _stft_dev(audio_f32, 1, T, self._cfg)
        out = torch.empty((7, T, 64), dtype=torch.float32, device=audio_f32.device)
        mean, istd = self.scaler_device
        st = (C.c_int64 * 4)(7 * T * 64, T * 64, 64, 1)
        gws = torch.empty(4, dtype=torch.int32, device=out.device)
        flags = torch.zeros(1, dtype=torch.int32, device=out.device)
        check(L.adyolo_logmel_from_stft(ptr(spec), 1, T, 4, 4, C.byref(self._cfg), ptr(mean), ptr(istd), ptr(out), st,
                                        ptr(gws), 1, stream_ptr()), "adyolo_logmel_from_stft")
        iv_out = out[4:]
        check(L.adyolo_iv_from_stft(ptr(spec), 1, T, C.byref(self._cfg), C.c_void_p(mean[4:].data_ptr()),
                                    C.c_void_p(istd[4:].data_ptr()), C.c_void_p(iv_out.data_ptr()), st, ptr(flags),
                                    stream_ptr()), "adyolo_iv_from_stft")
        if int(flags.item()) & 1:
            raise FloatingPointError("Feature extraction is generating nan outputs")
        return out

    def features_batched(self, audio_i16: torch.Tensor, out=None) -> torch.Tensor:
        """Batched training-path entry: (B, N, 4) int16 cuda -> (B, 7, T, 64) float32 cuda."""
        return features_batched(audio_i16, self.scaler_device, out=out, cfg=self._cfg)

    # ---- label half
    def get_yolo_label(self, label: dict, nb_label_frames: int):
        from .labels import get_yolo_label
        return get_yolo_label(label, nb_label_frames, self.grid)


# ------------------------------------------------------------------------------------------------
class FrontEndModule(torch.nn.Module):
    """The feature extractor as the first layer of the network: ``nn.Sequential(FrontEndModule(params), encoder)``.

    The reference extracts features inside ``Dataset.__getitem__`` (datasets.py:149-160) and its loops move the
    result to the device (train.py:48 ``feat.to(device).float()``).  With ``RawAudioDataset`` / ``collate_raw`` the
    batch that crosses that line is the int16 audio itself, (B, N, 4); ``.float()`` keeps every int16 value
    exactly, so this module casts back and runs the fused kernel: -> (B, 7, T, 64) float32, standardised with
    ``<data_pth>/scaler_wts.pkl`` (datasets.py:206,289-290).  In training mode SpecAug masks are drawn per clip and
    group with the reference's RNG sequence (augmentations.py:28-33) and applied on the device.  No parameters."""

    def __init__(self, params: dict, scaler: dict | None = None):
        super().__init__()
        from .augment import SpecAug
        self.proc = FeatureLabelProcessor(params, scaler=scaler)
        self.specaug = SpecAug(params, is_valid=False)

    def extract(self, audio_i16: torch.Tensor) -> torch.Tensor:
        return self.proc.features_batched(audio_i16)

    def forward(self, audio: torch.Tensor) -> torch.Tensor:
        if audio.dim() != 3 or audio.shape[-1] != 4:
            raise ValueError("FrontEndModule expects raw audio (B, N, 4)")
        if audio.dtype != torch.int16:
            audio = audio.to(torch.int16)              # exact: the values are int16 samples carried as float
        feat = self.extract(audio.contiguous())
        if self.training:
            feat = self.specaug.augment_batched(feat)
        return feat
