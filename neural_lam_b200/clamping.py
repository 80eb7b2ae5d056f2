"""Output clamping of the one-step predictors (reference models/step_predictors/base.py:181-396,
utils/tensor.py:7-81): instead of ``X_{t+1} = X_t + delta`` a variable with configured limits is updated as
``X_{t+1} = f(f^{-1}(X_t) + delta)`` with f = scaled sigmoid (both limits), shifted softplus (lower limit only) or
mirrored softplus (upper limit only); the limits are given in physical units and standardised with the state
mean / std.  Pure torch (elementwise, differentiable, device-agnostic); it only runs for models that configure
limits — the bench workloads have none and keep the fused step epilogue."""
import torch
from torch import nn


def inverse_softplus(x, beta=1.0, threshold=20.0):
    """Inverse of ``softplus(., beta)``; linear above ``threshold/beta``, input clamped away from 0
    (utils/tensor.py:7-50)."""
    # the lower bound is evaluated in float32 exactly as the reference does (log(float32(1 + 1e-6)) = 9.54e-7)
    lo = float(torch.log(torch.tensor(1e-6 + 1))) / beta
    xc = torch.clamp(x, min=lo, max=threshold / beta)
    non_linear = torch.log(torch.expm1(xc * beta)) / beta
    return torch.where(x * beta <= threshold, non_linear, x)


def inverse_sigmoid(x):
    """``log(x / (1-x))`` with x clamped into [1e-6, 1-1e-6] (utils/tensor.py:53-81)."""
    xc = torch.clamp(x, min=1e-6, max=1 - 1e-6)
    return torch.log(xc / (1 - xc))


class OutputClamp(nn.Module):
    """``lower`` / ``upper``: {state variable name: limit in physical units}; ``names``: the state variable names in
    feature order; ``state_mean`` / ``state_std``: (d_state,) standardisation statistics."""

    def __init__(self, names, lower, upper, state_mean, state_std):
        super().__init__()
        lower, upper = dict(lower or {}), dict(upper or {})
        unknown = (set(lower) | set(upper)) - set(names)
        if unknown:
            raise ValueError(f"State feature limits were provided for unknown features: {unknown}")
        both, lo_only, up_only = [], [], []
        both_lo, both_up, lo_l, up_l = [], [], [], []

        def norm(x, i):
            return (float(x) - float(state_mean[i])) / float(state_std[i])

        for i, n in enumerate(names):
            if n in lower and n in upper:
                if not lower[n] < upper[n]:
                    raise AssertionError(f'Invalid clamping limits for feature "{n}", lower: {lower[n]}, larger than upper: {upper[n]}')
                both.append(i)
                both_lo.append(norm(lower[n], i))
                both_up.append(norm(upper[n], i))
            elif n in lower:
                lo_only.append(i)
                lo_l.append(norm(lower[n], i))
            elif n in upper:
                up_only.append(i)
                up_l.append(norm(upper[n], i))
        f32, i64 = torch.float32, torch.long
        self.register_buffer("sigmoid_lower_lims", torch.tensor(both_lo, dtype=f32))
        self.register_buffer("sigmoid_upper_lims", torch.tensor(both_up, dtype=f32))
        self.register_buffer("softplus_lower_lims", torch.tensor(lo_l, dtype=f32))
        self.register_buffer("softplus_upper_lims", torch.tensor(up_l, dtype=f32))
        self.register_buffer("clamp_lower_upper_idx", torch.tensor(both, dtype=i64))
        self.register_buffer("clamp_lower_idx", torch.tensor(lo_only, dtype=i64))
        self.register_buffer("clamp_upper_idx", torch.tensor(up_only, dtype=i64))

    @property
    def active(self):
        return clamp_active(self)

    def forward(self, state_delta, prev_state):
        return clamped_update(self, state_delta, prev_state)


BUFFER_NAMES = ("sigmoid_lower_lims", "sigmoid_upper_lims", "softplus_lower_lims", "softplus_upper_lims",
                "clamp_lower_upper_idx", "clamp_lower_idx", "clamp_upper_idx")


def clamp_active(self):
    """``self``: any module holding the seven clamping buffers (an ``OutputClamp`` or, as in the reference where
    ``prepare_clamping_params`` registers them on the predictor itself, a step predictor)."""
    return (self.clamp_lower_upper_idx.numel() + self.clamp_lower_idx.numel() + self.clamp_upper_idx.numel()) > 0


def clamped_update(self, state_delta, prev_state):
        """``get_clamped_new_state`` (step_predictors/base.py:335-396); sharpness 1, centre 0 as in the reference."""
        new_state = prev_state + state_delta
        if self.clamp_lower_upper_idx.numel() > 0:
            idx = self.clamp_lower_upper_idx
            lo, up = self.sigmoid_lower_lims.to(new_state.dtype), self.sigmoid_upper_lims.to(new_state.dtype)
            z = inverse_sigmoid((prev_state[:, :, idx] - lo) / (up - lo)) + state_delta[:, :, idx]
            new_state[:, :, idx] = lo + (up - lo) * torch.sigmoid(z)
        if self.clamp_lower_idx.numel() > 0:
            idx = self.clamp_lower_idx
            lo = self.softplus_lower_lims.to(new_state.dtype)
            z = inverse_softplus(prev_state[:, :, idx] - lo) + state_delta[:, :, idx]
            new_state[:, :, idx] = lo + torch.nn.functional.softplus(z)
        if self.clamp_upper_idx.numel() > 0:
            idx = self.clamp_upper_idx
            up = self.softplus_upper_lims.to(new_state.dtype)
            z = -inverse_softplus(up - prev_state[:, :, idx]) + state_delta[:, :, idx]
            new_state[:, :, idx] = up - torch.nn.functional.softplus(-z)
        return new_state
