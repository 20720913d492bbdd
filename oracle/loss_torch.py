"""Torch oracle for the AD-YOLO responsibility assignment + loss (TEST INFRASTRUCTURE — not a
product path; only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may import it).

Restates ``/root/reference/src/models/loss.py:156-251`` (``ADYOLOloss``) op for op, with one
change that does not touch the arithmetic: label tensors are created on ``self.device``
(the reference allocates them on the CPU at loss.py:228-232 and therefore cannot run on CUDA
under torch >= 2).  Because every FP32 op here is the same eager torch op the
reference executes, running this class on ``cuda`` gives the bit-exact D / mask / argmin oracle
for the hand-written kernel ("torch on the same GPU"); running it on ``cpu``
is bit-identical to the unmodified reference class (pinned by ``oracle/make_golden.py`` ->
``tests/golden/loss_ref.npz`` and ``tests/test_oracle_loss.py``).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn


class ADYOLOlossOracle(object):
    def __init__(self, params: dict):
        # loss.py:157-180
        self.device = torch.device(params["args"]["device"])
        self.nb_classes = params["data_config"]["nb_classes"]
        self.grid_size = torch.Tensor(params["train_config"]["grid_size"])
        self.nb_anchors = params["train_config"]["nb_anchors"]
        nb_azi_grids = np.divmod(360, self.grid_size[0])
        nb_ele_grids = np.divmod(180, self.grid_size[1])
        nb_azi_grids = int(nb_azi_grids[0]) + int(nb_azi_grids[1] != 0)
        nb_ele_grids = int(nb_ele_grids[0]) + int(nb_ele_grids[1] != 0)
        self.nb_grids = torch.Tensor([nb_azi_grids, nb_ele_grids]).long()
        self.nb_predicts = self.nb_grids.prod().item() * self.nb_anchors
        self.grid_offset = torch.stack(torch.meshgrid(torch.arange(self.nb_grids[0].item()),
                                                      torch.arange(self.nb_grids[1].item()),
                                                      indexing="ij"), dim=-1)
        self.grid_offset = self.grid_offset * self.grid_size - torch.Tensor([180., 90.]) + (self.grid_size * 0.5)
        self.grid_size, self.grid_offset = self.grid_size.to(self.device), self.grid_offset.to(self.device)
        self.train_unify = params["train_config"]["train_unify"]
        self.g_overlap = params["train_config"]["g_overlap"]
        self.loss_gains = params["train_config"]["loss_gains"]
        self.bce_loss = nn.BCELoss(reduction="mean")

    def distance_between_polar_coordinates(self, output_coord, target_coord):
        # loss.py:182-187
        output_coord, target_coord = torch.deg2rad(output_coord), torch.deg2rad(target_coord)
        dist = (torch.sin(output_coord[..., 1]) * torch.sin(target_coord[..., 1]) +
                torch.cos(output_coord[..., 1]) * torch.cos(target_coord[..., 1]) *
                torch.cos(torch.abs(output_coord[..., 0] - target_coord[..., 0])))
        return torch.rad2deg(torch.acos(torch.clip(dist, -1 + 1e-7, 1 - 1e-7)))

    def decode(self, logit: torch.Tensor):
        # loss.py:193-213
        B, T, _ = logit.size()
        output = logit.reshape(B, T, self.nb_grids[0], self.nb_grids[1], self.nb_anchors, -1)
        output = torch.cat([output[..., :self.nb_classes + 1].sigmoid().clone(),
                            output[..., self.nb_classes + 1:].tanh().clone()], dim=-1)
        output[..., -2:] = output[..., -2:] * (0.5 + self.g_overlap)
        output[..., -2:] = output[..., -2:] * self.grid_size
        output[..., -2:] = output[..., -2:] + self.grid_offset[None, None, :, :, None]
        output[..., -1] = torch.clamp(output[..., -1].clone(), -90, 90)
        bi, ti, gi, gj, ai = torch.where(output[..., -2] >= 180.)
        output[bi, ti, gi, gj, ai, torch.ones_like(ai).long() * -2] = output[bi, ti, gi, gj, ai, torch.ones_like(ai).long() * -2] - 360.
        bi, ti, gi, gj, ai = torch.where(output[..., -2] < -180.)
        output[bi, ti, gi, gj, ai, torch.ones_like(ai).long() * -2] = output[bi, ti, gi, gj, ai, torch.ones_like(ai).long() * -2] + 360.
        return output

    def assign(self, logit: torch.Tensor, target: torch.Tensor):
        """loss.py:193-226 -> D (M,5) f32, masks (3,M,5) bool, argmin (M,) int64."""
        output = self.decode(logit)
        target = target.to(self.device)
        bi, ti, gi, gj = target[:, 0].long(), target[:, 1].long(), target[:, 2].long(), target[:, 3].long()
        D = self.distance_between_polar_coordinates(output[bi, ti, gi, gj][..., -2:],
                                                    target[:, None, -2:].repeat(1, self.nb_anchors, 1))
        amin = D.min(dim=1)[-1]
        masks = []
        for train_unify in self.train_unify:
            m = (D < train_unify)
            m[range(len(D)), amin] = True
            masks.append(m)
        return D, torch.stack(masks, 0), amin

    def __call__(self, logit: torch.Tensor, target: torch.Tensor):
        # loss.py:189-251
        B, T, _ = logit.size()
        output = self.decode(logit)
        target = target.to(self.device)
        bi, ti, gi, gj = target[:, 0].long(), target[:, 1].long(), target[:, 2].long(), target[:, 3].long()
        D = self.distance_between_polar_coordinates(output[bi, ti, gi, gj][..., -2:],
                                                    target[:, None, -2:].repeat(1, self.nb_anchors, 1))
        total_loss = torch.tensor([0.], device=self.device)
        dev = self.device
        for i, train_unify in enumerate(self.train_unify):
            responsible_mask = (D < train_unify)
            responsible_mask[range(len(D)), D.min(dim=1)[-1]] = True
            mi, ai = torch.where(responsible_mask)
            bi, ti, gi, gj, ci = (target[mi, 0].long(), target[mi, 1].long(), target[mi, 2].long(),
                                  target[mi, 3].long(), target[mi, 4].long())
            obj_label = torch.zeros(B, T, self.nb_grids[0], self.nb_grids[1], self.nb_anchors, device=dev).bool()
            obj_label[bi, ti, gi, gj, ai] = True
            cls_label = torch.zeros(B, T, self.nb_grids[0], self.nb_grids[1], self.nb_anchors, self.nb_classes, device=dev)
            cls_label[bi, ti, gi, gj, ai, ci] = 1.
            cls_label = cls_label[obj_label].to(self.device)
            class_loss = self.bce_loss(output[obj_label][..., 1:self.nb_classes + 1], cls_label)
            pos_object_loss = self.bce_loss(output[obj_label][..., 0], torch.ones(obj_label.sum().item(), device=self.device))
            neg_object_loss = self.bce_loss(output[~obj_label][..., 0], torch.zeros((~obj_label).sum().item(), device=self.device))
            if i == 0:
                total_loss = total_loss + (D[responsible_mask] / 180.).mean() * self.loss_gains["angular_gain"]
            total_loss = total_loss + (pos_object_loss * self.loss_gains["object_gain"] +
                                       neg_object_loss * self.loss_gains["nonobj_gain"] +
                                       class_loss * self.loss_gains["class_gain"]) / len(self.train_unify)
        return total_loss


def default_params(nb_classes=12, device="cpu"):
    """The slice of the reference's ``params`` dict the hot path reads (hyp_train.yaml:11-26,
    hyp_data_DCASE2021.yaml) — values, not code."""
    return {
        "args": {"device": device, "loss": "adyolo"},
        "data_config": {"nb_classes": nb_classes, "sr": 24000, "hop_length_s": 0.025, "win_length_s": 0.05,
                        "hop_length": 600, "win_length": 1200, "n_fft": 1200, "mel_bins": 64,
                        "window": "han", "label_hop_len_s": 0.1, "data_pth": None},
        "train_config": {"grid_size": [45, 45], "nb_anchors": 5, "g_overlap": 0.5,
                         "train_unify": [45., 25., 10.],
                         "loss_gains": {"angular_gain": 5., "object_gain": 1., "nonobj_gain": 5., "class_gain": 3.}},
    }
