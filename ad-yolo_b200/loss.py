"""AD-YOLO loss: Python mirror of ``models/loss.py:156-251`` over the CUDA C ABI.

``ADYOLOloss(params)(logit, target) -> tensor[1]`` like the reference, differentiable w.r.t.
``logit`` through a ``torch.autograd.Function`` whose forward launches the fused
assignment + loss + gradient kernels (no host synchronisation, unlike the reference's nine).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, ptr, require_cuda, stream_ptr
from .labels import DeviceRows, GridSpec

_ws_cache = {}


def _workspace(nbytes, device):
    key = (str(device), torch.cuda.current_stream(device).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def _check_inputs(logit, target, grid):
    require_cuda(logit, "ADYOLOloss")
    if logit.dim() != 3 or logit.shape[-1] != grid.nb_predicts * grid.nb_channels:
        raise ValueError(f"logit must be (B, T, {grid.nb_predicts * grid.nb_channels}); got {tuple(logit.shape)}")
    if target.dim() != 2 or target.shape[-1] != 7:
        raise ValueError("target must be (M, 7) [batch, frame, Gi, Gj, class, U, V]")


def adyolo_assign(logit: torch.Tensor, target: torch.Tensor, grid: GridSpec):
    """loss.py:193-226 only: -> D (M, A) float32, masks (n_thr, M, A) bool, argmin (M,) int64."""
    _check_inputs(logit, target, grid)
    logit = logit.detach().contiguous().float()
    target = target.to(logit.device, torch.float32).contiguous()
    B, T, _ = logit.shape
    M, A, K = target.shape[0], grid.nb_anchors, len(grid.train_unify)
    with torch.cuda.device(logit.device):
        D = torch.empty((M, A), dtype=torch.float32, device=logit.device)
        mask = torch.empty((K, M, A), dtype=torch.uint8, device=logit.device)
        amin = torch.empty((M,), dtype=torch.int32, device=logit.device)
        check(_lib.lib().adyolo_assign(ptr(logit), ptr(target), M, B, T, C.byref(grid.c), ptr(D), ptr(mask), ptr(amin),
                                       stream_ptr()), "adyolo_assign")
    return D, mask.bool(), amin.long()


class _ADYOLOFn(torch.autograd.Function):
    """Forward launches assignment + loss; when ``logit`` needs a gradient the same pass over the
    logits also writes d loss / d logit (the normalisers are known before that pass), so backward
    is a single launch that scales the stored gradient by ``grad_output`` on the device -- and
    returns without touching memory when that is exactly 1 (``loss.backward()``, train.py:54)."""

    @staticmethod
    def forward(ctx, logit, target, grid, n_rows_dev=None):
        B, T, _ = logit.shape
        M = target.shape[0]
        L = _lib.lib()
        need_grad = logit.requires_grad
        with torch.cuda.device(logit.device):
            loss = torch.empty(1, dtype=torch.float32, device=logit.device)
            nbytes = L.adyolo_loss_workspace_bytes(B, T, C.byref(grid.c))
            # the workspace carries label bits / counts to a repeated backward: private per call when needed
            ws = torch.empty(nbytes, dtype=torch.uint8, device=logit.device) if need_grad else _workspace(nbytes, logit.device)
            grad = torch.empty_like(logit) if need_grad else None
            if n_rows_dev is None:
                check(L.adyolo_loss(ptr(logit), ptr(target), M, B, T, C.byref(grid.c), ptr(loss), ptr(grad), None, None,
                                    None, ptr(ws), stream_ptr()), "adyolo_loss")
            else:
                check(L.adyolo_loss_devcount(ptr(logit), ptr(target), M, ptr(n_rows_dev), B, T, C.byref(grid.c),
                                             ptr(loss), ptr(grad), ptr(ws), stream_ptr()), "adyolo_loss_devcount")
        grid._last_ws = ws                  # ADYOLOloss.last_bad_rows() reads the skipped-row counter from it
        ctx.grid = grid
        ctx.grad = grad                     # consumed (scaled in place and handed out) by the first backward
        ctx.save_for_backward(logit, ws)
        return loss

    @staticmethod
    def backward(ctx, gout):
        logit, ws = ctx.saved_tensors
        B, T, _ = logit.shape
        gout = gout.to(torch.float32).contiguous()
        L = _lib.lib()
        with torch.cuda.device(logit.device):
            grad, ctx.grad = ctx.grad, None
            if grad is not None:
                check(L.adyolo_loss_grad_scale(ptr(grad), grad.numel(), ptr(gout), stream_ptr()), "adyolo_loss_grad_scale")
            else:                           # backward again (retain_graph=True): rebuild from the workspace
                grad = torch.empty_like(logit)
                check(L.adyolo_loss_backward(ptr(logit), B, T, C.byref(ctx.grid.c), ptr(ws), ptr(gout), ptr(grad),
                                             stream_ptr()), "adyolo_loss_backward")
        return grad, None, None, None


class ADYOLOloss(object):
    """loss.py:156-251.  Reads the same ``params`` keys as the reference."""

    def __init__(self, params: dict, check_rows: bool | None = None):
        """``check_rows`` (default: env ``ADYOLO_CHECK_ROWS=1`` or ``params['args']['check_rows']``): read the
        kernel's skipped-row counter back after every call (one host sync) and raise ``IndexError`` when a
        target row indexed outside the logit tensor -- what the reference's indexing does on such rows
        (loss.py:216,228-232).  Off by default to keep the step free of host syncs; ``last_bad_rows()``
        gives the same counter on demand (e.g. every N steps)."""
        import os
        if check_rows is None:
            check_rows = bool(params["args"].get("check_rows", False)) or os.environ.get("ADYOLO_CHECK_ROWS", "0") == "1"
        self.check_rows = bool(check_rows)
        self.device = torch.device(params["args"]["device"])
        if self.device.type != "cuda":
            raise RuntimeError("adyolo_b200.ADYOLOloss needs params['args']['device'] to be a CUDA device "
                               "(no CPU fallback)")
        tc = params["train_config"]
        self.nb_classes = params["data_config"]["nb_classes"]
        self.nb_anchors = tc["nb_anchors"]
        self.train_unify = tc["train_unify"]
        self.g_overlap = tc["g_overlap"]
        self.loss_gains = tc["loss_gains"]
        self.grid = GridSpec(self.nb_classes, self.nb_anchors, tc["grid_size"], tc["g_overlap"], tc["train_unify"],
                             tc["loss_gains"])
        self.nb_grids = torch.Tensor(self.grid.nb_grids).long()
        self.nb_predicts = self.grid.nb_predicts

    def distance_between_polar_coordinates(self, output_coord, target_coord):
        """loss.py:182-187 on already-decoded coordinates (debug/compat surface).  The fused path
        never materialises decoded outputs; this evaluates the same FP32 op sequence with torch ops
        on the caller's device."""
        output_coord, target_coord = torch.deg2rad(output_coord), torch.deg2rad(target_coord)
        dist = (torch.sin(output_coord[..., 1]) * torch.sin(target_coord[..., 1]) +
                torch.cos(output_coord[..., 1]) * torch.cos(target_coord[..., 1]) *
                torch.cos(torch.abs(output_coord[..., 0] - target_coord[..., 0])))
        return torch.rad2deg(torch.acos(torch.clip(dist, -1 + 1e-7, 1 - 1e-7)))

    def assign(self, logit, target):
        return adyolo_assign(logit, target, self.grid)

    def rows_from_events(self, events: torch.Tensor, nb_label_frames: int) -> DeviceRows:
        """(E, 5) event table [batch, frame, class, azi, ele] (what ``collate_raw`` hands over) -> target rows on the
        device (get_yolo_label + collate_fn, datasets.py:457-482,175-184), no host synchronisation."""
        from .labels import label_rows_batched
        ev = events.to(self.device, torch.float64)
        return label_rows_batched(ev, nb_label_frames, self.grid,
                                  max_rows=max(ev.shape[0], 1) * self.grid.nb_grids[0] * self.grid.nb_grids[1])

    def __call__(self, logit: torch.Tensor, target):
        n_rows = None
        if torch.is_tensor(target) and target.dim() == 2 and target.shape[-1] == 5:
            target = self.rows_from_events(target, logit.shape[1])          # raw-audio path: events, not rows
        if isinstance(target, DeviceRows):
            target, n_rows = target.rows, target.n_rows
        _check_inputs(logit, target, self.grid)
        if logit.dtype != torch.float32:
            logit = logit.float()
        logit = logit.contiguous()
        target = target.to(logit.device, torch.float32).contiguous()   # loss.py:199
        loss = _ADYOLOFn.apply(logit, target, self.grid, n_rows)
        if self.check_rows:
            bad = self.last_bad_rows()
            if bad:
                raise IndexError(f"ADYOLOloss: {bad} target rows index outside the logit tensor "
                                 f"(batch/frame/Gi/Gj/class out of range for logit {tuple(logit.shape)})")
        return loss

    def last_bad_rows(self) -> int:
        """Target rows the most recent call skipped as out of range (host sync)."""
        ws = getattr(self.grid, "_last_ws", None)
        if ws is None:
            return 0
        off = _lib.lib().adyolo_loss_bad_rows_offset()
        return int(ws[off:off + 4].view(torch.int32).item())


class WrapperCriterion(object):
    """wrapper.py:63-88 restricted to the loss on the B200 hot path."""

    def __init__(self, params):
        self.nb_classes = params["data_config"]["nb_classes"]
        self.loss_nm = params["args"]["loss"]
        if self.loss_nm == "adyolo":
            self.loss = ADYOLOloss(params)
        elif self.loss_nm in ("seddoa", "masked-seddoa", "accdoa", "adpit"):
            raise NotImplementedError(f"loss: {self.loss_nm} is outside the B200 hot path (use the reference)")
        else:
            raise NotImplementedError("loss: {}".format(self.loss_nm))

    def __call__(self, output, target):
        return self.loss(output, target)
