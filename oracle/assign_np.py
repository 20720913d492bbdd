"""CPU oracle for the AD-YOLO grid-cell label assignment (TEST INFRASTRUCTURE — not a product path).

Restates, in numpy float64 (exact arithmetic: every bound is a multiple of 22.5):

* ``/root/reference/src/datasets.py:219-238``  grid constants of ``FeatureLabelProcessor.__init__``
* ``/root/reference/src/datasets.py:457-482``  ``get_yolo_label``
* ``/root/reference/src/datasets.py:164-184``  ``collate_fn`` (label half)

Pinned against the unmodified reference functions by ``oracle/make_golden.py`` ->
``tests/golden/assign_cells.npz`` (checked in ``tests/test_oracle_assign.py``).
"""
from __future__ import annotations

import numpy as np


def grid_constants(grid_size=(45, 45), g_overlap=0.5):
    """datasets.py:219-235 -> (nb_grids [Ga, Ge], grid_offset, grid_lb, grid_ub), each (Ga,Ge,2) f64."""
    grid_size = np.array(grid_size)
    nb_azi = np.divmod(360, grid_size[0])
    nb_ele = np.divmod(180, grid_size[1])
    nb_azi = int(nb_azi[0]) + int(nb_azi[1] != 0)
    nb_ele = int(nb_ele[0]) + int(nb_ele[1] != 0)
    off = np.stack(np.meshgrid(np.arange(nb_azi), np.arange(nb_ele), indexing="ij"), axis=-1)
    off = off * grid_size - np.array([180, 90]) + (grid_size * 0.5)
    lb = off - (grid_size * (0.5 + g_overlap))
    lb[..., -1] = np.clip(lb[..., -1], -90, 90)
    ub = off + (grid_size * (0.5 + g_overlap))
    ub[..., -1] = np.clip(ub[..., -1], -90, 90)
    return [nb_azi, nb_ele], off, lb, ub


def get_yolo_label(label: dict, nb_label_frames: int, grid_size=(45, 45), g_overlap=0.5):
    """datasets.py:457-482.  Does NOT mutate ``label`` (the reference rewrites azi 180 -> -180
    in the caller's dict; the returned rows carry the rewritten value either way)."""
    _, _, lb, ub = grid_constants(grid_size, g_overlap)
    label_list = []
    for frame_idx, active_event_list in label.items():
        if frame_idx < nb_label_frames:
            for event in active_event_list:
                azi, ele = event[2], event[3]
                if azi == 180:
                    azi = -180.0
                azi_r = (lb[..., 0] <= azi) & (azi < ub[..., 0])
                ele_r = (lb[..., 1] <= ele) & (ele < ub[..., 1])
                resp = azi_r & ele_r
                resp |= (azi + 360 < ub[..., 0]) & ele_r
                resp |= (lb[..., 0] < azi - 360) & ele_r
                Gi, Gj = np.where(resp)
                for i, j in zip(Gi, Gj):
                    label_list.append([frame_idx, i, j, event[0], azi, ele])
    return label_list


def collate_labels(label_lists):
    """datasets.py:175-184 (label half): list over clips of label_list -> (M,7) float32
    rows [batch, frame, Gi, Gj, class, U, V].  Raises like the reference when all are empty."""
    rows = []
    for i, lst in enumerate(label_lists):
        if lst == []:
            continue
        a = np.asarray(lst, dtype=np.float32)
        rows.append(np.concatenate([np.full((len(lst), 1), i, np.float32), a], axis=-1))
    if not rows:
        raise ValueError("collate_fn: no events in batch (reference raises in torch.cat)")
    return np.concatenate(rows, 0)


def events_to_rows(events: np.ndarray, nb_label_frames, grid_size=(45, 45), g_overlap=0.5):
    """Array form used for the 10^6-frame stress config: ``events`` (E,5) float64
    [batch, frame, class, azi, ele] in dataset order -> (M,7) float64 rows, same per-event
    expansion and ordering as ``get_yolo_label`` + ``collate_fn``."""
    _, _, lb, ub = grid_constants(grid_size, g_overlap)
    lba, uba = lb[:, 0, 0], ub[:, 0, 0]
    lbe, ube = lb[0, :, 1], ub[0, :, 1]
    ev = np.asarray(events, dtype=np.float64)
    azi = np.where(ev[:, 3] == 180, -180.0, ev[:, 3])
    ele = ev[:, 4]
    a = azi[:, None]
    ar = ((lba[None] <= a) & (a < uba[None])) | (a + 360 < uba[None]) | (lba[None] < a - 360)
    er = (lbe[None] <= ele[:, None]) & (ele[:, None] < ube[None])
    ok = ev[:, 1] < nb_label_frames
    resp = ar[:, :, None] & er[:, None, :] & ok[:, None, None]
    e, gi, gj = np.nonzero(resp)  # row-major == event order, then (Gi,Gj) order
    return np.stack([ev[e, 0], ev[e, 1], gi.astype(np.float64), gj.astype(np.float64),
                     ev[e, 2], azi[e], ele[e]], axis=1)
