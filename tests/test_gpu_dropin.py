"""Executed drop-in check, device half (B200): the product classes -- RawAudioDataset + collate_raw (DataLoader
workers), FrontEndModule ahead of an encoder, WrapperCriterion taking the event table, LabelPostProcessor -- run
through the steps of the reference's train_one_epoch / test_epoch (src/train.py:44-60, src/test.py:33-60; the
reference itself is not on the GPU box, tests/test_dropin_reference_loops.py runs its unmodified loops on the
same wiring on the CPU).  Every step is checked against the CPU oracles: features of each batch, the loss of
each iteration on the same logits, and the SELD detections of each test clip."""
import os
import random

import numpy as np
import pytest
import torch

from conftest import ROOT  # noqa: F401
from oracle import assign_np, features_np as F
from oracle.loss_torch import ADYOLOlossOracle
from oracle.nms_torch import YoloPostOracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def A():
    import __graft_entry__ as g
    g.build()
    import adyolo_b200
    assert torch.cuda.is_available()
    return adyolo_b200


class TinyEncoder(torch.nn.Module):
    def __init__(self, nb_classes=12):
        super().__init__()
        self.pool = torch.nn.AvgPool2d((4, 8))
        self.lin = torch.nn.Linear(7 * 8, 160 * (nb_classes + 3))

    def forward(self, x):
        return self.lin(self.pool(x).permute(0, 2, 1, 3).flatten(2))


def test_quick_test_epoch_and_evaluation_with_the_product_classes(A, tmp_path, scaler2021):
    from _tiny_tree import build, params_for
    from torch.utils.data import DataLoader
    build(str(tmp_path), scaler2021)
    params = params_for(tmp_path, device="cuda:0")
    dev = torch.device("cuda:0")
    random.seed(7); np.random.seed(7); torch.manual_seed(7)
    ds = A.RawAudioDataset(params, "train")
    dl = DataLoader(ds, batch_size=2, shuffle=False, collate_fn=A.collate_raw, num_workers=2)
    front, enc = A.FrontEndModule(params), TinyEncoder().to(dev)
    model = torch.nn.Sequential(front, enc)
    crit = A.WrapperCriterion(params)
    oracle = ADYOLOlossOracle(params)                          # torch ops on the same GPU
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)

    # ---- train_one_epoch (train.py:44-60), --quick_test: 5 iterations
    model.train()
    losses = []
    for i, (feat, label) in enumerate(dl):
        assert feat.dtype == torch.int16 and label.shape[1] == 5
        raw = feat.clone()
        feat, label = feat.to(dev).float(), label.to(dev).float()         # train.py:48
        x = front(feat)
        ref = np.stack([F.features_foa_stack(a.numpy(), scaler=scaler2021) for a in raw])
        got = x.detach().cpu().numpy()
        assert (np.abs(got[:, :4] - ref[:, :4]) / np.maximum(np.abs(ref[:, :4]), 1.0)).max() < 1e-4
        assert np.abs(got[:, 4:] - ref[:, 4:]).max() < 1e-3
        output = enc(x)
        opt.zero_grad()
        loss = crit(output, label)
        rows = torch.from_numpy(assign_np.events_to_rows(label.cpu().numpy().astype(np.float64), output.shape[1]).astype(np.float32))
        want = oracle(output.detach(), rows.to(dev))
        assert abs(loss.item() - want.item()) <= 1e-5 * abs(want.item()), (loss.item(), want.item())
        loss.backward()
        opt.step()
        losses.append(loss.item())
        if params["args"]["quick_test"] and i == 4:
            break
    assert len(losses) == 5 and all(np.isfinite(losses))
    assert crit.loss.last_bad_rows() == 0

    # ---- test_epoch (test.py:33-60): batch 1, loss + post-processing per clip
    ds_t = A.RawAudioDataset(params, "test", is_valid=True)
    dl_t = DataLoader(ds_t, batch_size=1, shuffle=False, collate_fn=A.collate_raw)
    post = A.LabelPostProcessor(params)
    tc = params["train_config"]
    post_o = YoloPostOracle(12, conf_thresh=tc["conf_thresh"], clss_thresh=tc["clss_thresh"], unify_thresh=tc["unify_thresh"], device="cuda")
    model.eval()
    n_det = 0
    with torch.no_grad():
        for feat, label in dl_t:
            feat, label = feat.to(dev).float(), label.to(dev).float()
            output = model(feat)
            assert output.shape == (1, 30, 2400)
            assert np.isfinite(crit(output, label).item())
            seld = post.postprocess(output.detach().cpu())                # test.py:52
            want = post_o.clip_output(output[0])
            assert sorted(seld) == sorted(want)
            for fr in seld:
                a, b = np.asarray(seld[fr], np.float64), np.asarray(want[fr], np.float64)
                assert a.shape == b.shape and np.array_equal(a[:, 0], b[:, 0]) and np.abs(a - b).max() < 1e-5
                n_det += len(a)
    assert len(ds_t) == 2


def test_rotation_on_the_host_equals_the_fused_device_rotation(A, scaler2021):
    """RawAudioDataset rotates the int16 clip + events in the worker (augment.rotate_host); the fused form rotates
    inside the kernels.  Same features, same target rows, all 16 combinations."""
    from adyolo_b200.augment import rotate_host
    from adyolo_b200.features import _scaler_to_device
    rng = np.random.default_rng(5)
    clips = np.clip(rng.standard_normal((16, 24000, 4)) * 1500, -32767, 32767).astype(np.int16)
    ev = np.array([[b, t, (b + t) % 12, float(rng.integers(-180, 181)), float(rng.integers(-80, 81))] for b in range(16) for t in range(10)])
    sd = _scaler_to_device(scaler2021, ("MEL", "IV"), torch.device("cuda"))
    comb = torch.arange(16, dtype=torch.int8, device="cuda")
    grid = A.labels.GridSpec(12, 5, [45, 45], 0.5)
    fused = A.features_batched(torch.from_numpy(clips).cuda(), sd, rot_comb=comb)
    rows_fused = A.label_rows_batched(torch.from_numpy(ev).cuda(), 10, grid, rot_comb=comb)
    host_a, host_e = [], []
    for b in range(16):
        a, e = rotate_host(clips[b], ev[ev[:, 0] == b][:, 1:], b)
        host_a.append(a)
        host_e.append(np.concatenate([np.full((len(e), 1), float(b)), e], 1))
    plain = A.features_batched(torch.from_numpy(np.stack(host_a)).cuda(), sd)
    rows_host = A.label_rows_batched(torch.from_numpy(np.concatenate(host_e)).cuda(), 10, grid)
    assert torch.equal(rows_fused, rows_host)
    assert (fused - plain).abs().max().item() < 2e-4


def test_front_end_and_loss_on_two_streams_equal_the_single_stream_step(A):
    """bench.py issues the two independent halves of a step (front end | label rows + loss) on two CUDA streams.  Every
    entry point launches on torch's current stream and keeps its state (tables, launch counter, workspaces) per call or
    behind a mutex, so the results must be those of the same calls on one stream: features bit-identical, the label rows
    identical, loss and gradient identical up to the order of the floating-point atomics (<= 1e-6 relative)."""
    from oracle.loss_torch import default_params
    rng = np.random.default_rng(3)
    B, T, C = 24, 50, 12
    audio = torch.from_numpy(np.clip(rng.standard_normal((B, 120000, 4)) * 3000, -32768, 32767).astype(np.int16)).cuda()
    ev = np.array([[b, t, rng.integers(0, C), rng.integers(-180, 181), rng.integers(-90, 91)]
                   for b in range(B) for t in range(T) for _ in range(int(rng.integers(0, 3)))], dtype=np.float64)
    ev = torch.from_numpy(ev).cuda()
    grid = A.labels.GridSpec(C, 5, [45, 45], 0.5)
    crit = A.ADYOLOloss(default_params(C, "cuda:0"))
    logit = torch.randn((B, T, 2400), device="cuda", generator=torch.Generator(device="cuda").manual_seed(7)).requires_grad_(True)

    def loss_side():
        rows = A.label_rows_batched(ev, T, grid, max_rows=4 * ev.shape[0])
        logit.grad = None
        loss = crit(logit, rows)
        loss.backward()
        return rows.materialize() if hasattr(rows, "materialize") else rows, loss.detach().clone(), logit.grad.clone()

    f0 = A.features_batched(audio, None)
    r0, l0, g0 = loss_side()
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for _ in range(4):                                   # several overlapping steps in flight
        with torch.cuda.stream(s1):
            f = A.features_batched(audio, None)
        with torch.cuda.stream(s2):
            outs.append((f,) + loss_side())
    torch.cuda.synchronize()
    for f, r, l, g in outs:
        assert torch.equal(f, f0)
        assert torch.equal(r, r0)
        assert abs(l.item() - l0.item()) <= 1e-6 * abs(l0.item())
        assert (g - g0).abs().max().item() <= 1e-6 * g0.abs().max().item() + 1e-12
