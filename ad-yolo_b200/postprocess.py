"""AD-YOLO post-processing (decode + conn-merge NMS): Python mirror of
``datasets.py:485-533,741-857`` (``LabelPostProcessor`` restricted to ``--loss adyolo``) over the
CUDA C ABI.  ``postprocess(output)`` keeps the reference signature (batch-1 logits in, dict
``{frame: [[class, x, y, z], ...]}`` out); ``yolo_post_batched`` is the tensor form."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, ptr, require_cuda, stream_ptr
from .labels import GridSpec


NMS_MODES = {"conn-merge": 0, "soft-merge": 1}       # any other string -> plain NMS (datasets.py:834)


def nms_mode(name: str) -> int:
    return NMS_MODES.get(name, 2)


def yolo_post_batched(logit: torch.Tensor, grid: GridSpec, conf_thresh: float, clss_thresh: float, unify_thresh: float,
                      max_det: int | None = None, nms: str = "conn-merge"):
    """logit (B, T, Ga*Ge*A*(C+3)) float32 CUDA -> det (B, T, max_det, 4) float32 [class, x, y, z],
    count (B, T) int32.  ``max_det`` defaults to the exact upper bound (anchors x classes per frame);
    with a smaller cap the call raises if any frame produced more detections."""
    require_cuda(logit, "yolo_post_batched")
    if logit.dim() != 3 or logit.shape[-1] != grid.nb_predicts * grid.nb_channels:
        raise ValueError(f"logit must be (B, T, {grid.nb_predicts * grid.nb_channels})")
    logit = logit.detach().contiguous().float()
    B, T, _ = logit.shape
    if max_det is None:
        max_det = grid.nb_predicts * grid.nb_classes
    with torch.cuda.device(logit.device):
        det = torch.zeros((B, T, max_det, 4), dtype=torch.float32, device=logit.device)
        count = torch.zeros((B, T), dtype=torch.int32, device=logit.device)
        over = torch.zeros(1, dtype=torch.int32, device=logit.device)
        check(_lib.lib().adyolo_yolo_post(ptr(logit), B * T, C.byref(grid.c), float(conf_thresh), float(clss_thresh),
                                          float(unify_thresh), nms_mode(nms), int(max_det), ptr(det), ptr(count), ptr(over), stream_ptr()),
              "adyolo_yolo_post")
    if int(over.item()):
        raise RuntimeError(f"yolo_post_batched: a frame produced more than max_det={max_det} detections")
    return det, count


class LabelPostProcessor:
    """datasets.py:485-533 for ``params['args']['loss'] == 'adyolo'``."""

    def __init__(self, params):
        self.nb_classes = params["data_config"]["nb_classes"]
        self.loss = params["args"]["loss"]
        if self.loss != "adyolo":
            raise NotImplementedError("postprocess: {} (only adyolo is on the B200 path)".format(self.loss))
        tc = params["train_config"]
        self.conf_thresh = tc["conf_thresh"]
        self.clss_thresh = tc["clss_thresh"]
        self.unify_thresh = tc["unify_thresh"]
        self.g_overlap = tc["g_overlap"]
        self.nms = tc["nms"]                          # 'conn-merge' | 'soft-merge' | anything else = plain NMS
        self.nb_anchors = tc["nb_anchors"]
        self.grid = GridSpec(self.nb_classes, tc["nb_anchors"], tc["grid_size"], tc["g_overlap"],
                             tc.get("train_unify", [45., 25., 10.]), tc.get("loss_gains"))
        self.postprocess = self.get_yolo_output

    def get_conf_thresh(self):
        return self.conf_thresh

    def set_conf_thresh(self, thresh):
        self.conf_thresh = thresh
        self.clss_thresh = thresh

    def get_yolo_output(self, batch_yolo_output: torch.Tensor, max_det: int | None = None):
        """datasets.py:741-857: (1, T, 160*(C+3)) logits -> {frame_idx: [[class_idx, X, Y, Z], ...]}."""
        if batch_yolo_output.dim() != 3 or batch_yolo_output.shape[0] != 1:
            raise ValueError("get_yolo_output expects (1, T, n) logits (the reference evaluates with batch 1)")
        dev = batch_yolo_output.device if batch_yolo_output.is_cuda else torch.device("cuda")
        det, count = yolo_post_batched(batch_yolo_output.to(dev), self.grid, self.conf_thresh, self.clss_thresh,
                                       self.unify_thresh, max_det, self.nms)
        det, count = det[0].cpu(), count[0].cpu().tolist()
        return {t: det[t, :n].tolist() for t, n in enumerate(count) if n > 0}
