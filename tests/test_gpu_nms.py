"""GPU decode + conn-merge NMS vs the reference LabelPostProcessor (CPU golden)
and vs the torch restatement executed on the same GPU."""
import numpy as np
import pytest
import torch

from oracle.loss_torch import default_params
from oracle.nms_torch import YoloPostOracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def A():
    import __graft_entry__ as g
    g.build()
    import adyolo_b200
    return adyolo_b200


def _params(C=12):
    p = default_params(C, "cuda:0")
    p["train_config"].update({"conf_thresh": 0.5, "clss_thresh": 0.5, "unify_thresh": 15., "nms": "conn-merge"})
    return p


def _compare(got: dict, want: dict, tol=2e-6):
    assert set(got) == set(want)
    for fr in want:
        g, w = np.asarray(got[fr]), np.asarray(want[fr])
        assert g.shape == w.shape, fr
        assert np.array_equal(g[:, 0], w[:, 0]), fr                     # classes, order
        assert np.abs(g[:, 1:] - w[:, 1:]).max() < tol, fr


def test_postprocess_vs_reference_golden(A, gold):
    g = gold("nms_ref.npz")
    post = A.LabelPostProcessor(_params())
    got = post.postprocess(torch.from_numpy(g["logit"]).cuda())
    want = {}
    for r in g["rows"]:
        want.setdefault(int(r[0]), []).append(list(r[1:]))
    assert len(want) > 0 and max(len(v) for v in want.values()) >= 2
    _compare(got, want)


@pytest.mark.parametrize("C,scale", [(12, 1.5), (13, 2.5), (14, 0.7)])
def test_postprocess_vs_torch_oracle_on_gpu(A, C, scale):
    gen = torch.Generator(device="cuda").manual_seed(C)
    T = 12
    logit = torch.randn((2, T, 160 * (C + 3)), device="cuda", generator=gen) * scale
    y = logit.view(2, T, 8, 4, 5, C + 3)
    y[:, :, 1, 2, :, 0] += 4; y[:, :, 1, 2, :, 4] += 4; y[:, ::3, 2, 2, :3, 0] += 4; y[:, ::3, 2, 2, :3, 4] += 4
    post = A.LabelPostProcessor(_params(C))
    orc = YoloPostOracle(nb_classes=C, device="cuda")
    for thr in ((0.5, 0.3) if C == 12 else (0.5,)):
        post.set_conf_thresh(thr)
        orc.conf_thresh = orc.clss_thresh = thr
        for b in range(2):
            _compare(post.postprocess(logit[b:b + 1]), orc.clip_output(logit[b]))


def test_overflow_and_modes(A):
    post = A.LabelPostProcessor(_params())
    big = torch.full((1, 2, 2400), 6.0, device="cuda")
    with pytest.raises(RuntimeError):
        post.get_yolo_output(big, max_det=4)
    p = _params(); p["train_config"]["nms"] = "soft-merge"
    with pytest.raises(NotImplementedError):
        A.LabelPostProcessor(p)
