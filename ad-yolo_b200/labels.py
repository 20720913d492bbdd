"""AD-YOLO grid-cell label assignment: Python mirror over the CUDA C ABI.

Reference surface mirrored:
  * ``datasets.py:219-238``  grid constants           -> ``GridSpec``
  * ``datasets.py:457-482``  ``get_yolo_label``        -> ``get_yolo_label`` (dict in, list out)
  * ``datasets.py:164-184``  ``collate_fn``            -> ``collate_fn``
plus the array form used on the device-resident training path: ``label_rows_batched``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import GridCfg, check, ptr, require_cuda, stream_ptr


class GridSpec:
    """train_config slice shared by the label builder and the loss (grid_size, nb_anchors,
    g_overlap, train_unify, loss_gains) + data_config.nb_classes."""

    def __init__(self, nb_classes, nb_anchors, grid_size, g_overlap, train_unify=(45., 25., 10.), loss_gains=None):
        loss_gains = loss_gains or {"angular_gain": 5., "object_gain": 1., "nonobj_gain": 5., "class_gain": 3.}
        if len(train_unify) > 4:
            raise NotImplementedError("at most 4 train_unify thresholds")
        gs = [float(grid_size[0]), float(grid_size[1])]
        na = np.divmod(360, gs[0]); ne = np.divmod(180, gs[1])
        self.nb_grids = [int(na[0]) + int(na[1] != 0), int(ne[0]) + int(ne[1] != 0)]   # datasets.py:222-226
        self.nb_classes, self.nb_anchors = int(nb_classes), int(nb_anchors)
        self.grid_size, self.g_overlap = gs, float(g_overlap)
        self.train_unify = [float(t) for t in train_unify]
        self.loss_gains = dict(loss_gains)
        tu = (C.c_float * 4)(*(self.train_unify + [0.0] * (4 - len(self.train_unify))))
        self.c = GridCfg(self.nb_classes, self.nb_anchors, (C.c_double * 2)(*gs), self.g_overlap,
                         len(self.train_unify), tu, float(loss_gains["angular_gain"]),
                         float(loss_gains["object_gain"]), float(loss_gains["nonobj_gain"]),
                         float(loss_gains["class_gain"]))

    @property
    def nb_predicts(self):
        return self.nb_grids[0] * self.nb_grids[1] * self.nb_anchors

    @property
    def nb_channels(self):
        return self.nb_classes + 3


class DeviceRows:
    """Target rows whose count is only known on the device: ``rows`` has capacity rows and the first
    ``n_rows`` (int64 tensor of shape (1,) on the device) are valid.  ``ADYOLOloss`` accepts it in
    place of the (M, 7) tensor, so label rows and loss are enqueued with no host synchronisation."""

    def __init__(self, rows: torch.Tensor, n_rows: torch.Tensor):
        self.rows, self.n_rows = rows, n_rows

    def overflowed(self) -> bool:
        """True when more rows were produced than ``rows`` can hold (one host sync)."""
        return int(self.n_rows.item()) > self.rows.shape[0]

    def materialize(self) -> torch.Tensor:
        n = int(self.n_rows.item())
        if n > self.rows.shape[0]:
            raise RuntimeError(f"label rows overflow: {n} rows produced, capacity {self.rows.shape[0]} "
                               "(size max_rows as E * Ga * Ge; 4 * E only holds for g_overlap <= 0.5)")
        return self.rows[:n]


def label_rows_batched(events: torch.Tensor, nb_label_frames: int, grid: GridSpec, return_cellmask=False,
                       max_rows: int | None = None, rot_comb: torch.Tensor | None = None):
    """events (E, 5) float64 on CUDA, rows [batch, frame, class, azi, ele] in dataset order
    -> target rows (M, 7) float32 on CUDA [batch, frame, Gi, Gj, class, U, V]
    (== get_yolo_label per clip followed by collate_fn's label half).

    With ``max_rows`` (``E * Ga * Ge`` always suffices; ``4 * E`` does for the reference's g_overlap = 0.5)
    the result is a ``DeviceRows`` and nothing synchronises with the host; rows beyond the capacity are
    dropped and ``DeviceRows.overflowed()`` / ``materialize()`` report it.
    ``rot_comb`` (int8 CUDA tensor, one RotationAug combination 0..15 per clip / batch index): the
    label half of the rotation augmentation is applied before the cell test."""
    require_cuda(events, "label_rows_batched")
    if events.dtype != torch.float64 or events.dim() != 2 or events.shape[1] != 5:
        raise ValueError("events must be a float64 tensor of shape (E, 5)")
    events = events.contiguous()
    E = events.shape[0]
    L = _lib.lib()
    with torch.cuda.device(events.device):
        ws = torch.empty(max(L.adyolo_label_workspace_bytes(E), 64), dtype=torch.uint8, device=events.device)
        cellmask = torch.empty(max(E, 1), dtype=torch.int32, device=events.device)
        total = torch.empty(1, dtype=torch.int64, device=events.device)   # always written by adyolo_label_cells
        if rot_comb is not None:
            if rot_comb.dtype != torch.int8 or not rot_comb.is_cuda:
                raise ValueError("rot_comb must be an int8 CUDA tensor")
            rot_comb = rot_comb.contiguous()
        n_rot = 0 if rot_comb is None else rot_comb.numel()
        if max_rows is not None:
            rows = torch.empty((int(max_rows), 7), dtype=torch.float32, device=events.device)
            check(L.adyolo_label_cells_rows(ptr(events), E, int(nb_label_frames), C.byref(grid.c), ptr(rot_comb), n_rot,
                                            ptr(cellmask), ptr(total), ptr(ws), ptr(rows), int(max_rows), stream_ptr()),
                  "adyolo_label_cells_rows")
            return DeviceRows(rows, total)
        check(L.adyolo_label_cells(ptr(events), E, int(nb_label_frames), C.byref(grid.c), ptr(rot_comb), n_rot,
                                   ptr(cellmask), ptr(total), ptr(ws), stream_ptr()), "adyolo_label_cells")
        M = int(total.item())   # the one host sync of the label path (sizes the output)
        rows = torch.empty((M, 7), dtype=torch.float32, device=events.device)
        check(L.adyolo_label_rows(ptr(events), E, C.byref(grid.c), ptr(rot_comb), n_rot, ptr(cellmask), ptr(ws),
                                  ptr(rows), M, stream_ptr()), "adyolo_label_rows")
    if return_cellmask:
        return rows, cellmask[:E]
    return rows


def get_yolo_label(label: dict, nb_label_frames: int, grid: GridSpec):
    """datasets.py:457-482.  label {frame: [[cls, src, azi, ele], ...]} -> list of
    [frame_idx, Gi, Gj, class_idx, U, V] (python numbers, reference ordering).  The caller's dict
    is not modified."""
    require_cuda(None, "get_yolo_label")
    ev = [[0.0, float(fr), float(e[0]), float(e[2]), float(e[3])] for fr, evs in label.items() for e in evs]
    if not ev:
        return []
    events = torch.tensor(ev, dtype=torch.float64, device="cuda")
    rows, cm = label_rows_batched(events, nb_label_frames, grid, return_cellmask=True)
    cm = cm.cpu().numpy().astype(np.uint32)
    # rebuild exact python values (float64 azi/ele, int frame/class) in the kernel's row order
    out = []
    Ge = grid.nb_grids[1]
    for e, m in zip(ev, cm):
        m = int(m)
        azi = -180.0 if e[3] == 180 else e[3]
        cell = 0
        while m:
            if m & 1:
                out.append([int(e[1]), cell // Ge, cell % Ge, int(e[2]), azi, e[4]])
            m >>= 1
            cell += 1
    assert len(out) == rows.shape[0]
    return out


def collate_fn(batch):
    """datasets.py:164-184: (feature (C,T,F), label_list) items -> (B,C,T,F), (M,7) float32
    [batch, frame, Gi, Gj, class, U, V].  Raises (like torch.cat in the reference) when every clip
    is event-free."""
    feat, label = zip(*batch)
    batch_label_list = []
    for i, label_list in enumerate(label):
        if label_list == []:
            continue
        batch_label_list.append(
            torch.cat([torch.Tensor([i] * len(label_list)).unsqueeze(-1), torch.Tensor(label_list)], dim=-1))
    return torch.stack(feat, 0), torch.cat(batch_label_list, 0)
