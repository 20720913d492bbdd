"""Pin the torch loss oracle against the unmodified reference ADYOLOloss (CPU golden)."""
import numpy as np
import pytest
import torch

from oracle.loss_torch import ADYOLOlossOracle, default_params


@pytest.mark.parametrize("C", [12, 13, 14])
def test_oracle_bitexact_vs_reference_cpu(gold, C):
    g = gold("loss_ref.npz")
    logit = torch.from_numpy(g[f"C{C}_logit"]).requires_grad_(True)
    target = torch.from_numpy(g[f"C{C}_target"])
    orc = ADYOLOlossOracle(default_params(C, "cpu"))
    loss = orc(logit, target)
    loss.backward()
    np.testing.assert_array_equal(loss.detach().numpy(), g[f"C{C}_loss"])
    np.testing.assert_array_equal(logit.grad.numpy(), g[f"C{C}_grad"])
    D, masks, amin = orc.assign(logit.detach(), target)
    np.testing.assert_array_equal(D.numpy(), g[f"C{C}_D"])
    assert masks.shape == (3, len(target), 5) and masks.dtype == torch.bool
    assert masks[:, torch.arange(len(target)), amin].all()


def test_f9_threshold_cases_present(gold):
    """The golden targets contain the pole / threshold elevations of SURVEY F9."""
    g = gold("loss_ref.npz")
    el = np.abs(g["C12_target"][:, 6])
    assert {65.0, 45.0, 80.0, 90.0} <= set(np.unique(el))


def test_loss_nan_when_no_targets():
    orc = ADYOLOlossOracle(default_params(12, "cpu"))
    out = orc(torch.zeros(1, 2, 2400), torch.zeros(0, 7))
    assert out.shape == (1,) and torch.isnan(out).all()
