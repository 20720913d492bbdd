"""Raw-audio data path either side of the hot path.  Host plumbing only.

* ``load_wav2npy`` / ``load_csv2dict``: the reference's on-disk formats (``utility.py:219-246``,
  ``datasets.py:99-118``): int16 4-channel wav, 5-column polar csv ``frame,class,source,azi,ele``.
* ``chunk_plan`` / ``ResidentClips``: the ``chunking`` action (``preprocess.py:13-85``) without the
  41x-duplicated chunk files: whole training files stay resident in HBM as int16, a chunk is a
  *view* (sample offset) handed to ``adyolo_features_foa_views`` and its label rows are the file's
  events windowed and re-indexed exactly as ``chunk_instance`` does.  Chunk names follow the
  reference's ``<file>_chunkNNN`` convention so samplers and logs stay comparable.
* ``EpochSampler``: the without-replacement epoch sampling of ``datasets.py:67-92`` (same
  ``random`` call sequence, so the same seed yields the same file lists; the pool of remaining
  files is checkpointable like ``get_remaining_file`` / ``init_remaining_file_from_list``).
"""
from __future__ import annotations

import copy
import random

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, require_cuda, stream_ptr
import ctypes as C


def load_wav2npy(wav_pth):
    """datasets.py:99-101 / utility.py:219-221 -> int16 (T, 4)."""
    import scipy.io.wavfile as wav
    _, audio = wav.read(wav_pth)
    return audio


def load_csv2dict(csv_pth):
    """The reference's metadata format (datasets.py:103-118 / utility.py:233-246): one event per line,
    ``frame,class,source,azi,ele`` (polar) or ``frame,class,source,x,y,z`` (Cartesian) ->
    {frame: [[class, source, coords...], ...]} in file order.  A frame whose line has another
    column count still gets its (possibly empty) list, as in the reference."""
    table: dict[int, list] = {}
    with open(csv_pth, "r") as fh:
        for raw in fh:
            cols = raw.strip().split(",")
            events = table.setdefault(int(cols[0]), [])
            if len(cols) in (5, 6):
                events.append([int(cols[1]), int(cols[2])] + [float(c) for c in cols[3:]])
    return table


def chunk_plan(n_samples: int, sr=24000, chunk_window_s=20, chunk_stride_s=1, label_hop_len_s=0.1):
    """Geometry of ``chunk_instance`` (preprocess.py:27-37) for one file of ``n_samples`` samples:
    -> dict(pad, n_chunks, wav_window, wav_stride, csv_window, csv_stride)."""
    wav_window = sr * chunk_window_s
    wav_stride = sr * chunk_stride_s
    csv_window = int(chunk_window_s / label_hop_len_s)
    csv_stride = int(chunk_stride_s / label_hop_len_s)
    if n_samples < wav_window:
        raise ValueError("file shorter than one chunk window")   # the reference's sliding_window_view raises too
    rem = (n_samples - wav_window) % wav_stride
    pad = wav_stride - rem if rem != 0 else 0
    n_chunks = (n_samples + pad - wav_window) // wav_stride + 1
    return {"pad": int(pad), "n_chunks": int(n_chunks), "wav_window": int(wav_window), "wav_stride": int(wav_stride),
            "csv_window": csv_window, "csv_stride": csv_stride}


class ResidentClips:
    """Training files resident on the device as one int16 (S, 4) buffer + their event tables."""

    def __init__(self, device="cuda", sr=24000, chunk_window_s=20, chunk_stride_s=1, label_hop_len_s=0.1):
        self.device = torch.device(device)
        self.geom = dict(sr=sr, chunk_window_s=chunk_window_s, chunk_stride_s=chunk_stride_s,
                         label_hop_len_s=label_hop_len_s)
        self._host, self._files, self._events, self._plans, self._index = [], [], [], [], {}
        self._cursor = 0
        self.audio = None

    def add(self, name: str, audio_i16: np.ndarray, label: dict):
        a = np.ascontiguousarray(audio_i16, dtype=np.int16)
        if a.ndim != 2 or a.shape[1] != 4:
            raise ValueError("audio must be int16 (T, 4)")
        plan = chunk_plan(len(a), **self.geom)
        if plan["pad"]:
            a = np.pad(a, [(0, plan["pad"]), (0, 0)], "constant")      # preprocess.py:33
        ev = np.asarray([[fr, e[0], e[2], e[3]] for fr, evs in label.items() for e in evs], dtype=np.float64).reshape(-1, 4)
        self._index[name] = len(self._files)
        self._files.append((name, self._cursor, len(a)))
        self._events.append(ev)                                         # [frame, class, azi, ele] in dict order
        self._plans.append(plan)
        self._host.append(a)
        self._cursor += len(a)
        self.audio = None

    def finalize(self):
        if self.device.type == "cuda":
            require_cuda(None, "ResidentClips")
        host = np.concatenate(self._host, 0) if self._host else np.zeros((0, 4), np.int16)
        self.audio = torch.from_numpy(host).to(self.device)
        return self

    def chunk_names(self):
        """Names the reference's chunking action would have written (preprocess.py:80)."""
        return [f"{name}_chunk{i + 1:03d}" for (name, _, _), p in zip(self._files, self._plans) for i in range(p["n_chunks"])]

    def _resolve(self, chunk_name: str):
        base, _, num = chunk_name.rpartition("_chunk")
        fi = self._index[base]
        return fi, int(num) - 1

    def batch(self, chunk_names):
        """-> (offsets int64 (B,) on device, events (E,5) float64 on device [batch, frame, class, azi, ele],
        N samples per chunk, nb_label_frames).  Event order = chunk order, then the file's dict order
        restricted to the chunk window (== load_csv2dict of the chunk csv written by write_dict2csv)."""
        offs, evs = [], []
        for b, cn in enumerate(chunk_names):
            fi, ci = self._resolve(cn)
            _, start, _ = self._files[fi]
            p = self._plans[fi]
            if not 0 <= ci < p["n_chunks"]:
                raise IndexError(cn)
            offs.append(start + ci * p["wav_stride"])        # in range by construction: ci < n_chunks of the padded file
            ev = self._events[fi]
            f0 = ci * p["csv_stride"]
            sel = (ev[:, 0] >= f0) & (ev[:, 0] < f0 + p["csv_window"])
            e = ev[sel]
            # chunk_instance walks frame_idx = 0..window-1 in order (preprocess.py:41-45)
            order = np.argsort(e[:, 0], kind="stable")
            e = e[order]
            evs.append(np.column_stack([np.full(len(e), b, np.float64), e[:, 0] - f0, e[:, 1], e[:, 2], e[:, 3]]))
        p0 = self._plans[self._resolve(chunk_names[0])[0]]
        events = np.concatenate(evs, 0) if evs else np.zeros((0, 5))
        return (torch.tensor(offs, dtype=torch.int64, device=self.device),
                torch.from_numpy(events.reshape(-1, 5)).to(self.device), p0["wav_window"], p0["csv_window"])


def features_batched_views(audio_resident: torch.Tensor, offsets, n_samples: int, scaler_dev=None,
                           rot_comb: torch.Tensor | None = None, apply_topdb: bool = True,
                           validate: bool | None = None) -> torch.Tensor:
    """Features of B clip *views* of a resident int16 (S, 4) buffer -> (B, 7, T, 64) float32.

    ``offsets``: first sample of each view.  A python sequence / CPU tensor is range-checked on the
    host before upload (no device synchronisation); a CUDA tensor (e.g. from ``ResidentClips.batch``,
    which has already checked its own values on the host) is trusted unless ``validate=True``, which
    costs one host sync."""
    from .features import _cfg, _workspace
    require_cuda(audio_resident, "features_batched_views")
    if audio_resident.dtype != torch.int16 or audio_resident.dim() != 2 or audio_resident.shape[1] != 4:
        raise ValueError("resident audio must be an int16 tensor of shape (S, 4)")
    S = audio_resident.shape[0]
    if not torch.is_tensor(offsets):
        offsets = torch.as_tensor(list(offsets), dtype=torch.int64)
    if validate is None:
        validate = not offsets.is_cuda
    if validate and offsets.numel():
        lo, hi = (int(v) for v in torch.stack([offsets.min(), offsets.max()]).tolist())
        if lo < 0 or hi + n_samples > S:
            raise IndexError("clip view out of range")
    offsets = offsets.to(audio_resident.device, torch.int64).contiguous()
    B = offsets.shape[0]
    if rot_comb is not None:
        if rot_comb.dtype != torch.int8 or rot_comb.shape != (B,) or not rot_comb.is_cuda:
            raise ValueError("rot_comb must be an int8 CUDA tensor of shape (B,)")
        rot_comb = rot_comb.contiguous()
    cfg = _cfg()
    L = _lib.lib()
    T = n_samples // 600
    with torch.cuda.device(audio_resident.device):
        out = torch.empty((B, 7, T, 64), dtype=torch.float32, device=audio_resident.device)
        ws = _workspace(L.adyolo_frontend_workspace_bytes(C.byref(cfg), B, n_samples), audio_resident.device)
        mean, istd = scaler_dev if scaler_dev is not None else (None, None)
        check(L.adyolo_features_foa_views(ptr(audio_resident.contiguous()), ptr(offsets), B, n_samples, C.byref(cfg),
                                          ptr(mean), ptr(istd), ptr(rot_comb), ptr(out), ptr(ws),
                                          1 if apply_topdb else 0, stream_ptr()), "adyolo_features_foa_views")
    return out


def _take(pool: list, picks: list) -> list:
    """``pool`` minus one occurrence of every element of ``picks``, order preserved (what a
    sequence of ``list.remove`` calls leaves behind)."""
    budget: dict = {}
    for name in picks:
        budget[name] = budget.get(name, 0) + 1
    kept = []
    for name in pool:
        if budget.get(name, 0):
            budget[name] -= 1
        else:
            kept.append(name)
    return kept


class RawAudioDataset(torch.utils.data.Dataset):
    """Drop-in for the reference ``datasets.Dataset`` (datasets.py:21-162) on the B200 path: same constructor
    (``params, set_type, is_valid``), same directory layout, file listing, epoch sampling and checkpoint hooks,
    same attributes the orchestration reads (``loss_nm``, ``get_filelist()``) -- but ``__getitem__`` stops before
    the arithmetic: DataLoader workers cannot use CUDA, so an item is the raw material of the hot path,

        (audio int16 (N, 4) tensor, events float64 (E, 4) tensor [frame, class, azi, ele])

    and ``collate_raw`` stacks it.  Features, label rows and the loss then run on the device
    (``FrontEndModule`` in front of the encoder, ``ADYOLOloss`` accepting the event table), which is what lets the
    reference's own ``train_one_epoch`` / ``test_epoch`` (train.py:44-60, test.py:33-60) run unchanged.

    Rotation augmentation: applied here, on the int16 clip and the event table, with the reference's draw
    (``int(random.uniform(0, 16))``, augmentations.py:93) -- per-item host work of a few hundred microseconds; the
    fused device variant (``features_batched(rot_comb=)``) is for pipelines that own their training loop."""

    def __init__(self, params: dict, set_type: str, is_valid: bool = False):
        import os
        from .augment import RotationAug
        dc, self.params = params["data_config"], params
        self.is_valid, self.is_infer, self.set_type = is_valid, set_type == "infer", set_type
        self.loss_nm = params["args"]["loss"]
        if self.loss_nm != "adyolo":
            raise NotImplementedError("loss: {} (only --loss adyolo is on the B200 hot path)".format(self.loss_nm))
        opj = os.path.join
        if set_type == "train":                                   # datasets.py:36-49
            sub = "dev-train-chunked_{}s_{}s".format(dc["chunk_window_s"], dc["chunk_stride_s"])
            self.wav_pth, self.csv_pth = opj(dc["data_pth"], "foa_dev", sub), opj(dc["data_pth"], "metadata_dev", sub)
            names = [f.replace(".wav", "") for f in os.listdir(self.wav_pth)]
            self.sampler = EpochSampler(names, params["train_config"]["batch_size"] * params["train_config"]["nb_iters"])
            self.sampler.sample_filelist_for_train_iter()
        else:                                                     # datasets.py:50-58
            if self.is_infer:
                self.wav_pth, self.csv_pth = str(params["args"]["infer_pth"]), None
            else:
                self.wav_pth = opj(dc["data_pth"], "foa_dev", "dev-{}".format(set_type))
                self.csv_pth = opj(dc["data_pth"], "metadata_dev", "dev-{}".format(set_type))
            self.sampler = None
            self._filelist = [f.replace(".wav", "") for f in os.listdir(self.wav_pth)]
        self.rotation = RotationAug(params, is_valid)

    # ---- the orchestration's hooks (train.py:150,176,247; test.py:36)
    @property
    def filelist(self):
        return self.sampler.filelist if self.sampler is not None else self._filelist

    def sample_filelist_for_train_iter(self):
        return self.sampler.sample_filelist_for_train_iter()

    def init_remaining_file_from_list(self, remaining_file):
        self.sampler.init_remaining_file_from_list(remaining_file)

    def get_remaining_file(self):
        return self.sampler.get_remaining_file()

    def get_filelist(self):
        return self.filelist

    def __len__(self):
        return len(self.filelist)

    def __getitem__(self, index):
        import os
        name = self.filelist[index]
        audio = np.ascontiguousarray(load_wav2npy(os.path.join(self.wav_pth, name + ".wav")), dtype=np.int16)
        label = {} if self.is_infer else load_csv2dict(os.path.join(self.csv_pth, name + ".csv"))
        ev = np.asarray([[fr, e[0], e[2], e[3]] for fr, evs in label.items() for e in evs], dtype=np.float64).reshape(-1, 4)
        if self.rotation.apply_augment:
            from .augment import rotate_host
            audio, ev = rotate_host(audio, ev, int(random.uniform(0, 16)))
        return torch.from_numpy(audio), torch.from_numpy(ev)


def collate_raw(batch):
    """collate_fn of the raw-audio path (mirror of datasets.py:164-184): items (audio (N,4) int16, events (E,4)) ->
    (audio (B, N, 4) int16, events (E_total, 5) float32 [batch, frame, class, azi, ele]).  float32 is exact for the
    integer-degree DCASE labels and survives the ``label.to(device).float()`` of train.py:48.  Like the reference
    (torch.cat of an empty list, datasets.py:184) it raises when no clip of the batch has an event."""
    audio, events = zip(*batch)
    rows = [torch.cat([torch.full((len(e), 1), float(i), dtype=torch.float64), e], dim=1) for i, e in enumerate(events) if len(e)]
    if not rows:
        raise RuntimeError("collate_raw: no events in the batch (the reference's collate_fn raises here too)")
    return torch.stack(audio, 0), torch.cat(rows, 0).to(torch.float32)


class EpochSampler:
    """Without-replacement epoch sampling of the reference Dataset (datasets.py:67-98): every epoch
    draws ``nb_samples`` names from a pool that survives across epochs (and checkpoints); when the
    pool cannot cover an epoch, what is left is used up first and the rest comes from a fresh copy
    of the full list.  The sequence of ``random.sample`` / ``random.shuffle`` calls and their
    arguments is the reference's, so a seeded ``random`` reproduces its file lists
    (golden ``chunking.npz::sampler_epochs``)."""

    def __init__(self, total_filelist, nb_samples: int):
        self.total_filelist = list(total_filelist)
        self.remaining_file = list(self.total_filelist)
        self.nb_samples = int(nb_samples)
        self.filelist: list = []

    def sample_filelist_for_train_iter(self):
        leftovers: list = []
        pool = self.remaining_file
        if len(pool) < self.nb_samples:
            if pool:                                   # use up the tail of the old pool, in shuffled order
                random.shuffle(pool)
                leftovers = list(pool)
            pool = list(self.total_filelist)           # and top up from a fresh pool
        drawn = random.sample(pool, self.nb_samples - len(leftovers))
        self.remaining_file = _take(pool, drawn)
        self.filelist = drawn + leftovers
        return self.filelist

    # checkpoint hooks (train.py:150,247)
    def init_remaining_file_from_list(self, remaining_file):
        self.remaining_file = remaining_file

    def get_remaining_file(self):
        return self.remaining_file

    def get_filelist(self):
        return self.filelist
