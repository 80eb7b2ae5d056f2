"""Output clamping (reference step_predictors/base.py:181-396, utils/tensor.py:7-81): the product module
``neural_lam_b200.clamping`` against (1) golden vectors generated from the reference's own ``utils/tensor.py``
(oracle/gen_golden_clamp.py), (2) the oracle restatement of ``get_clamped_new_state`` on seeded inputs, (3) the
properties the construction guarantees (limits respected, zero increment = identity inside the limits)."""
import os

import numpy as np
import pytest
import torch

from neural_lam_b200 import clamping
from oracle import reference_port as rp

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "clamp_inverse.npz")


def test_inverse_functions_match_reference_golden():
    g = np.load(GOLD)
    xs, xp = torch.from_numpy(g["softplus_x"]), torch.from_numpy(g["sigmoid_x"])
    torch.testing.assert_close(clamping.inverse_sigmoid(xp), torch.from_numpy(g["inv_sigmoid"]), rtol=0, atol=0)
    torch.testing.assert_close(rp.ref_inverse_sigmoid(xp), torch.from_numpy(g["inv_sigmoid"]), rtol=0, atol=0)
    for beta in (1.0, 2.5):
        want = torch.from_numpy(g[f"inv_softplus_beta{beta}"])
        torch.testing.assert_close(clamping.inverse_softplus(xs, beta=beta), want, rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(rp.ref_inverse_softplus(xs, beta=beta), want, rtol=0, atol=0)


NAMES = [f"v{i}" for i in range(6)]
LOWER = {"v0": 0.0, "v1": -3.0, "v4": 1.5}     # v0: both limits, v1: lower only, v4: both
UPPER = {"v0": 1.0, "v2": 7.0, "v4": 4.0}      # v2: upper only; v3, v5: unclamped


def _stats(dtype):
    g = torch.Generator().manual_seed(3)
    return torch.randn(6, generator=g, dtype=dtype), (torch.rand(6, generator=g, dtype=dtype) + 0.5)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_clamped_update_matches_oracle(dtype):
    mean, std = _stats(dtype)
    clamp = clamping.OutputClamp(NAMES, LOWER, UPPER, mean, std).to(dtype)
    assert clamp.active
    g = torch.Generator().manual_seed(0)
    prev = torch.randn(2, 50, 6, generator=g, dtype=dtype)
    # put the clamped variables inside their (standardised) limits, as states produced by the clamp are
    for i, n in enumerate(NAMES):
        lo = (LOWER[n] - mean[i]) / std[i] if n in LOWER else None
        up = (UPPER[n] - mean[i]) / std[i] if n in UPPER else None
        if lo is not None and up is not None:
            prev[:, :, i] = lo + (up - lo) * torch.rand(2, 50, generator=g, dtype=dtype)
        elif lo is not None:
            prev[:, :, i] = lo + torch.rand(2, 50, generator=g, dtype=dtype) * 3
        elif up is not None:
            prev[:, :, i] = up - torch.rand(2, 50, generator=g, dtype=dtype) * 3
    delta = torch.randn(2, 50, 6, generator=g, dtype=dtype)
    want = rp.clamped_new_state(delta, prev, NAMES, LOWER, UPPER, mean, std)
    got = clamp(delta, prev)
    # the limits are float32 buffers, as in the reference (torch.tensor of float32 statistics): ~1e-7 relative
    tol = 5e-6
    torch.testing.assert_close(got, want, rtol=tol, atol=tol)
    # unclamped variables: plain residual update
    torch.testing.assert_close(got[:, :, [3, 5]], (prev + delta)[:, :, [3, 5]], rtol=0, atol=0)
    # limits hold for any increment
    big = clamp(50 * delta, prev)
    for i, n in enumerate(NAMES):
        if n in LOWER:
            assert (big[:, :, i] >= (LOWER[n] - mean[i]) / std[i] - 1e-6).all()
        if n in UPPER:
            assert (big[:, :, i] <= (UPPER[n] - mean[i]) / std[i] + 1e-6).all()
    # zero increment: identity inside the limits
    same = clamp(torch.zeros_like(delta), prev)
    torch.testing.assert_close(same, prev, rtol=1e-4, atol=1e-4)
    # differentiable
    d = delta.clone().requires_grad_(True)
    clamp(d, prev).sum().backward()
    assert torch.isfinite(d.grad).all() and (d.grad[:, :, 3] == 1).all()


def test_clamp_configuration_errors_and_inactive():
    mean, std = _stats(torch.float32)
    with pytest.raises(ValueError, match="unknown features"):
        clamping.OutputClamp(NAMES, {"nope": 0.0}, {}, mean, std)
    with pytest.raises(AssertionError, match="Invalid clamping limits"):
        clamping.OutputClamp(NAMES, {"v0": 2.0}, {"v0": 1.0}, mean, std)
    none = clamping.OutputClamp(NAMES, {}, {}, mean, std)
    assert not none.active
    x, d = torch.randn(1, 4, 6), torch.randn(1, 4, 6)
    torch.testing.assert_close(none(d, x), x + d, rtol=0, atol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("with_boundary", [False, True])
def test_fused_clamped_step_epilogue_kernel(with_boundary):
    """nlam_step_epilogue_clamped (rescale + clamped update + boundary mix in one launch) vs the oracle restatement of
    get_clamped_new_state + the boundary mix of autoregressive.py:128-131; fp32 elementwise: 5e-6."""
    from neural_lam_b200 import models, ops, synthetic

    spec = synthetic.make_graph_spec(16, 16, hierarchical=False, n_levels=1)
    ds = synthetic.SyntheticDatastore(spec, d_state=6, d_forcing=6, d_static=1, boundary_width=1)
    ds.state_var_names = NAMES
    ds.state_mean, ds.state_std = _stats(torch.float32)
    model = models.GraphLAM(ds, spec, hidden_dim=16, processor_layers=1, output_clamping_lower=LOWER,
                            output_clamping_upper=UPPER).cuda()
    assert model.clamps_output and model._clamp_kind.tolist() == [1, 2, 3, 0, 1, 0]
    g = torch.Generator().manual_seed(0)
    B, G, D = 3, 256, 6
    prev = torch.randn(B, G, D, generator=g)
    for i, n in enumerate(NAMES):  # clamped variables start inside their limits
        lo = (LOWER[n] - ds.state_mean[i]) / ds.state_std[i] if n in LOWER else None
        up = (UPPER[n] - ds.state_mean[i]) / ds.state_std[i] if n in UPPER else None
        if lo is not None and up is not None:
            prev[:, :, i] = lo + (up - lo) * torch.rand(B, G, generator=g)
        elif lo is not None:
            prev[:, :, i] = lo + torch.rand(B, G, generator=g) * 3
        elif up is not None:
            prev[:, :, i] = up - torch.rand(B, G, generator=g) * 3
    net_out = torch.randn(B, G, D, generator=g) * 2
    std, mean = torch.rand(D, generator=g) + 0.5, torch.randn(D, generator=g) * 0.1
    bnd, mask = torch.randn(B, G, D, generator=g), (torch.rand(G, 1, generator=g) < 0.3).float()
    want = rp.clamped_new_state(net_out * std + mean, prev, NAMES, LOWER, UPPER, ds.state_mean, ds.state_std)
    if with_boundary:
        want = mask * bnd + (1 - mask) * want
    got = ops.step_epilogue(net_out.cuda(), prev.cuda(), bnd.cuda() if with_boundary else None,
                            mask.cuda() if with_boundary else None, std.cuda(), mean.cuda(),
                            clamp=(model._clamp_kind, model._clamp_lo, model._clamp_up))
    torch.testing.assert_close(got.cpu(), want, rtol=5e-6, atol=5e-6)
