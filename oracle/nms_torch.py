"""Torch oracle for the AD-YOLO decode + NMS post-processing (TEST INFRASTRUCTURE).

Restates ``/root/reference/src/datasets.py:741-919`` (``LabelPostProcessor.get_yolo_output`` with
its three ``nms`` branches -- 'conn-merge', 'soft-merge', plain -- and its helper functions) op for op, generalised to a batch of clips and
returning, per (clip, frame), the list ``[[class_idx, x, y, z], ...]`` the reference builds.
Pinned against the unmodified reference class by ``oracle/make_golden.py`` ->
``tests/golden/nms_ref.npz``.
"""
from __future__ import annotations

import copy

import numpy as np
import torch


def distance_between_polar_coordinates(coord1, coord2):
    # datasets.py:863-876
    coord1, coord2 = torch.deg2rad(coord1), torch.deg2rad(coord2)
    dist = torch.sin(coord1[..., 1]) * torch.sin(coord2[..., 1]) + torch.cos(coord1[..., 1]) * torch.cos(coord2[..., 1]) * torch.cos(torch.abs(coord1[..., 0] - coord2[..., 0]))
    return torch.rad2deg(torch.acos(torch.clip(dist, -1, 1)))


def _polar_to_cartesian(class_output):
    # datasets.py:879-896
    polar = torch.deg2rad(class_output[..., -2:])
    x = torch.cos(polar[..., [0]]) * torch.cos(polar[..., [1]])
    y = torch.sin(polar[..., [0]]) * torch.cos(polar[..., [1]])
    z = torch.sin(polar[..., [1]])
    return torch.cat([class_output[..., [0]], x, y, z], dim=-1)


def _voted(unifying_output, conf_thresh):
    # datasets.py:899-919
    polar = torch.deg2rad(unifying_output[..., -2:])
    x = torch.cos(polar[..., [0]]) * torch.cos(polar[..., [1]])
    y = torch.sin(polar[..., [0]]) * torch.cos(polar[..., [1]])
    z = torch.sin(polar[..., [1]])
    cart = torch.cat([x, y, z], dim=-1)
    w = torch.exp(unifying_output[..., 1] ** 2 / conf_thresh).softmax(dim=-1).unsqueeze(-1)
    v = (cart * w).sum(dim=0, keepdim=True)
    v = v / torch.sqrt((v ** 2).sum())
    return torch.cat([unifying_output[:1, [0]], v], dim=-1)


class YoloPostOracle:
    def __init__(self, nb_classes=12, grid_size=(45, 45), nb_anchors=5, g_overlap=0.5, conf_thresh=0.5,
                 clss_thresh=0.5, unify_thresh=15.0, device="cpu", nms="conn-merge"):
        self.nms = nms
        self.nb_classes, self.nb_anchors, self.g_overlap = nb_classes, nb_anchors, g_overlap
        self.conf_thresh, self.clss_thresh, self.unify_thresh = conf_thresh, clss_thresh, unify_thresh
        self.device = torch.device(device)
        gs = torch.Tensor(list(grid_size))
        na, ne = np.divmod(360, gs[0]), np.divmod(180, gs[1])
        self.nb_grids = [int(na[0]) + int(na[1] != 0), int(ne[0]) + int(ne[1] != 0)]
        off = torch.stack(torch.meshgrid(torch.arange(self.nb_grids[0]), torch.arange(self.nb_grids[1]), indexing="ij"), dim=-1)
        self.grid_offset = (off * gs - torch.Tensor([180., 90.]) + gs * 0.5).to(self.device)
        self.grid_size = gs.to(self.device)

    def decode(self, logit_clip):
        # datasets.py:752-770 for one clip (T, 160*(C+3))
        T = logit_clip.shape[0]
        y = logit_clip.reshape(T, self.nb_grids[0], self.nb_grids[1], self.nb_anchors, -1)
        y = torch.cat([y[..., :self.nb_classes + 1].sigmoid(), y[..., self.nb_classes + 1:].tanh()], dim=-1)
        y[..., -2:] = y[..., -2:] * (0.5 + self.g_overlap)
        y[..., -2:] = y[..., -2:] * self.grid_size
        y[..., -2:] = y[..., -2:] + self.grid_offset[None, :, :, None]
        y[..., -1] = torch.clamp(y[..., -1], -90, 90 - 1e-7)
        t, gi, gj, a = torch.where(y[..., -2] >= 180.)
        y[t, gi, gj, a, torch.ones_like(a).long() * -2] = y[t, gi, gj, a, torch.ones_like(a).long() * -2] - 360.
        t, gi, gj, a = torch.where(y[..., -2] < -180.)
        y[t, gi, gj, a, torch.ones_like(a).long() * -2] = y[t, gi, gj, a, torch.ones_like(a).long() * -2] + 360.
        y[..., 1:self.nb_classes + 1] *= y[..., [0]]
        return y

    def clip_output(self, logit_clip):
        """-> {frame: [[cls, x, y, z], ...]} exactly as get_yolo_output builds it (conn-merge)."""
        y = self.decode(logit_clip.to(self.device))
        frames = [fo[torch.where(fo[..., 0] > self.conf_thresh)] for fo in y]
        out = {}
        for frame_cnt, fo in enumerate(frames):
            if len(fo) == 0:
                continue
            i, j = (fo[..., 1:self.nb_classes + 1] > self.clss_thresh).nonzero().t()
            fo = torch.cat([j.float().unsqueeze(1), fo[..., 1:self.nb_classes + 1][i, j].unsqueeze(1), fo[..., -2:][i]], dim=1)
            fo = fo[fo[..., 1].argsort(descending=True)]
            det = []
            for class_idx in fo[..., 0].unique():
                co = fo[fo[..., 0] == class_idx]
                if len(co) == 1:
                    det.append(_polar_to_cartesian(co))
                    continue
                if self.nms == "soft-merge":               # datasets.py:817-831
                    ref_out = copy.deepcopy(co)
                    while co.shape[0]:
                        D = distance_between_polar_coordinates(co[:1, -2:], ref_out[:, -2:])
                        det.append(_voted(ref_out[D <= self.unify_thresh], self.clss_thresh))
                        if len(co) == 1:
                            break
                        D = distance_between_polar_coordinates(co[:1, -2:], co[1:, -2:])
                        co = co[1:][D > self.unify_thresh]
                    continue
                if self.nms != "conn-merge":               # datasets.py:834-846 plain NMS
                    while co.shape[0]:
                        det.append(_polar_to_cartesian(co[:1]))
                        if len(co) == 1:
                            break
                        D = distance_between_polar_coordinates(co[:1, -2:], co[1:, -2:])
                        co = co[1:][D > self.unify_thresh]
                    continue
                D = distance_between_polar_coordinates(co[None, :, -2:].repeat(len(co), 1, 1), co[:, None, -2:].repeat(1, len(co), 1))
                ref = (D < self.unify_thresh)
                while co.shape[0]:
                    pre = torch.zeros(len(co), device=co.device).bool()
                    cur = copy.deepcopy(ref[0])
                    while not (pre == cur).all():
                        if cur.sum() == 1:
                            break
                        pre = copy.deepcopy(cur)
                        cur |= ref[cur].sum(dim=0).bool()
                    det.append(_voted(co[cur], self.clss_thresh))
                    co = co[~cur]
                    ref = ref[~cur][:, ~cur]
            if len(det):
                out[frame_cnt] = torch.cat(det, dim=0).tolist()
        return out
