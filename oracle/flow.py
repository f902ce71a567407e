"""ORACLE (test infrastructure only) -- flow matcher, ODE solver and LitModule numerics, fp32 PyTorch.

PARITY UNPINNED for the third-party parts (see oracle/unet.py header): restates
  * torchcfm 1.0.7 `ConditionalFlowMatcher.sample_location_and_conditional_flow`   
  * torchdyn 1.0.6 `NeuralODE(...).trajectory(x, t_span)` fixed-step + dopri5       
and follows, line by line, the in-repo numerics of
  * src/models/conditional_flow_matching.py:53-74   (model_step)
  * src/models/conditional_flow_matching.py:133-170 (generate)
  * src/models/class_conditional_flow_matching.py:49-71, 130-190
  * the mask / ROI variants (pinned bit-for-bit against the reference's own modules, tests/golden/mask_variants_small.pt):
    src/models/conditional_flow_matching_masked.py:59-92, conditional_flow_matching_ROI_loss.py:64-97,
    conditional_flow_matching_conditional_mask.py:54-82, 143-199, conditional_flow_matching_conditional_toggle_mask.py:68-87, 186-187
"""
from __future__ import annotations

import torch
import torch.nn as nn


def pad_t_like_x(t, x):
    if isinstance(t, (float, int)):
        return t
    return t.reshape(-1, *([1] * (x.dim() - 1)))


class ConditionalFlowMatcher:
    """torchcfm.conditional_flow_matching.ConditionalFlowMatcher (B.1)."""

    def __init__(self, sigma: float = 0.0):
        self.sigma = sigma

    def compute_mu_t(self, x0, x1, t):
        t = pad_t_like_x(t, x0)
        return t * x1 + (1 - t) * x0

    def compute_sigma_t(self, t):
        return self.sigma

    def sample_xt(self, x0, x1, t, epsilon):
        return self.compute_mu_t(x0, x1, t) + self.compute_sigma_t(t) * epsilon

    def compute_conditional_flow(self, x0, x1, t, xt):
        return x1 - x0

    def sample_location_and_conditional_flow(self, x0, x1, t=None, return_noise=False):
        if t is None:
            t = torch.rand(x0.shape[0]).type_as(x0)  # CPU default generator, then moved (finding 9)
        assert len(t) == x0.shape[0], "t has to have batch size dimension"
        eps = torch.randn_like(x0)
        xt = self.sample_xt(x0, x1, t, eps)
        ut = self.compute_conditional_flow(x0, x1, t, xt)
        if return_noise:
            return t, xt, ut, eps
        return t, xt, ut


# ----------------------------------------------------------------------------------------------- solver

_DOPRI_C = (0.0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0)
_DOPRI_A = (
    (),
    (1 / 5,),
    (3 / 40, 9 / 40),
    (44 / 45, -56 / 15, 32 / 9),
    (19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729),
    (9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656),
    (35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84),
)
_DOPRI_B5 = (35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0.0)
_DOPRI_B4 = (5179 / 57600, 0.0, 7571 / 16695, 393 / 640, -92097 / 339200, 187 / 2100, 1 / 40)


def _hairer_norm(x):
    return x.abs().pow(2).mean().sqrt()


def _init_step(f, f0, x0, t0, order, atol, rtol):
    scale = atol + x0.abs() * rtol
    d0, d1 = _hairer_norm(x0 / scale), _hairer_norm(f0 / scale)
    h0 = torch.tensor(1e-6, dtype=x0.dtype, device=x0.device) if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    x1 = x0 + h0 * f0
    f1 = f(t0 + h0, x1)
    d2 = _hairer_norm((f1 - f0) / scale) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = torch.max(torch.tensor(1e-6, dtype=x0.dtype, device=x0.device), h0 * 1e-3)
    else:
        h1 = (0.01 / max(float(d1), float(d2))) ** (1.0 / float(order + 1))
        h1 = torch.as_tensor(h1, dtype=x0.dtype, device=x0.device)
    return torch.min(100 * h0, h1)


def odeint(f, x, t_span, solver="dopri5", atol=1e-4, rtol=1e-4):
    """Returns (t_eval, sol[len(t_span), *x.shape]).  `f(t, x)` with 0-dim tensor t (B.2)."""
    t_span = t_span.to(x)
    sol = [x]
    if solver in ("euler", "midpoint", "rk4"):
        for k in range(len(t_span) - 1):
            t, dt = t_span[k], t_span[k + 1] - t_span[k]
            if solver == "euler":
                x = x + dt * f(t, x)
            elif solver == "midpoint":
                x = x + dt * f(t + 0.5 * dt, x + 0.5 * dt * f(t, x))
            else:  # torchdyn's rk4 is the 3/8-rule variant
                k1 = f(t, x)
                k2 = f(t + dt / 3, x + dt * k1 / 3)
                k3 = f(t + dt * 2 / 3, x + dt * (k2 - k1 / 3))
                k4 = f(t + dt, x + dt * (k1 - k2 + k3))
                x = x + dt * (k1 + 3 * (k2 + k3) + k4) / 8
            sol.append(x)
        return t_span, torch.stack(sol)
    if solver != "dopri5":
        raise NotImplementedError(solver)
    # adaptive dopri5 as torchdyn 1.0.6 runs it (`_adaptive_odeint`, `init_step`, `adapt_step`; restated from the published
    # source, B.2): initial step from Hairer's heuristic with exponent 1/(order+1), order = 5; a step that would pass the
    # next t_span point is cut to land on it (no dense output) and afterwards the controller continues from the REMAINDER
    # of the un-cut proposal, (dt_old - dt) * factor; factor = clamp(safety * ratio^(-1/order), min_factor, max_factor) with
    # min_factor raised to 1 for accepted steps (ratio < 1) and the plain max_factor for ratio == 0.
    t = t_span[0]
    k1 = f(t, x)
    order, safety, min_f, max_f = 5, 0.9, 0.2, 10.0
    dt = _init_step(f, k1, x, t, order, atol, rtol)
    ckpt = 1
    n_fe = 2
    while ckpt < len(t_span):
        t_target = t_span[ckpt]
        cut = bool(t + dt > t_target)
        dt_old = dt
        if cut:
            dt = t_target - t
        ks = [k1]
        for s in range(1, 7):
            xs = x + dt * sum(a * k for a, k in zip(_DOPRI_A[s], ks))
            ks.append(f(t + _DOPRI_C[s] * dt, xs))
        n_fe += 6
        x_new = x + dt * sum(b * k for b, k in zip(_DOPRI_B5, ks))
        x_err = dt * sum((b5 - b4) * k for b5, b4, k in zip(_DOPRI_B5, _DOPRI_B4, ks))
        err_ratio = float(_hairer_norm(x_err / (atol + rtol * torch.max(x.abs(), x_new.abs()))))
        if err_ratio <= 1.0:
            t = t_target if cut else t + dt
            x = x_new
            k1 = ks[-1]  # FSAL
            if cut:
                sol.append(x)
                ckpt += 1
        if cut:
            rest = dt_old - dt
            dt = rest if float(rest) > 0.0 else dt  # (guard: a remainder rounded to 0 would stall the loop)
        if err_ratio == 0.0:
            factor = max_f
        else:
            lo = 1.0 if err_ratio < 1.0 else min_f
            factor = min(max_f, max(lo, safety * err_ratio ** (-1.0 / order)))
        dt = dt * factor
    odeint.last_nfe = n_fe
    return t_span, torch.stack(sol)


class NeuralODE(nn.Module):
    """torchdyn.core.NeuralODE, restricted to what generate() uses (B.2)."""

    def __init__(self, vector_field, solver="tsit5", order=1, atol=1e-3, rtol=1e-3, sensitivity="autograd",
                 solver_adjoint=None, atol_adjoint=1e-4, rtol_adjoint=1e-4, interpolator=None, integral_loss=None,
                 seminorm=False, return_t_eval=True, optimizable_params=()):
        super().__init__()
        self.vf = vector_field
        self.solver, self.atol, self.rtol = solver, atol, rtol

    def trajectory(self, x, t_span):
        _, sol = odeint(lambda t, z: self.vf(t, z), x, t_span, solver=self.solver, atol=self.atol, rtol=self.rtol)
        return sol


# ----------------------------------------------------------------------------------------- LitModule numerics

def model_step(net, flow_matcher, batch, t=None):
    """src/models/conditional_flow_matching.py:53-74 and class_conditional_flow_matching.py:49-71."""
    x0, x1 = batch[:2]
    y = batch[2].long() if (len(batch) > 2 and getattr(net, "num_classes", None) is not None) else None
    t, xt, ut = flow_matcher.sample_location_and_conditional_flow(x0, x1, t=t)
    vt = net(t, xt) if y is None else net(t, xt, y=y)
    return torch.mean((vt - ut) ** 2)


def model_step_mask_weighted(net, flow_matcher, batch, t=None, roi_lambda=10.0):
    """src/models/conditional_flow_matching_masked.py:59-92: MSE weighted by 1 + roi_lambda * mask."""
    x0, x1, mask = batch
    t, xt, ut = flow_matcher.sample_location_and_conditional_flow(x0, x1, t=t)
    vt = net(t, xt)
    weights = (1.0 + roi_lambda * mask).expand_as(vt)
    return (weights * (vt - ut) ** 2).sum() / (weights.sum() + 1e-8)


def model_step_roi(net, flow_matcher, batch, t=None, lambda_roi=1.0):
    """src/models/conditional_flow_matching_ROI_loss.py:64-97: MSE + lambda_roi * ROI Charbonnier(xt, x1)."""
    x0, x1, mask = batch
    t, xt, ut = flow_matcher.sample_location_and_conditional_flow(x0, x1, t=t)
    vt = net(t, xt)
    loss_fm = torch.mean((vt - ut) ** 2)
    m = mask.float()
    diff = xt - x1
    charb = torch.sqrt(diff * diff + 1e-3 * 1e-3)
    roi_charb = (charb * m).sum() / (m.sum() * xt.shape[1] + 1e-8)
    return loss_fm + lambda_roi * roi_charb


def model_step_mask_conditioned(net, flow_matcher, batch, t=None, use_mask_toggle=False):
    """src/models/conditional_flow_matching_conditional_mask.py:68-82 (and ..._toggle_mask.py:68-87 with the toggle):
    the mask is the UNet's 4th input channel."""
    x0, x1, mask = batch
    if use_mask_toggle and torch.rand(1).item() < 0.5:
        mask = torch.zeros_like(mask)
    t, xt, ut = flow_matcher.sample_location_and_conditional_flow(x0, x1, t=t)
    vt = net(t, torch.cat([xt, mask], dim=1))
    return torch.mean((vt - ut) ** 2)


@torch.no_grad()
def generate_mask_conditioned(net, source_img, mask, num_steps=100, solver="dopri5", atol=1e-4, rtol=1e-4,
                              zero_mask=False):
    """conditional_flow_matching_conditional_mask.py:143-199 (zero_mask: ..._toggle_mask.py:186-187)."""
    net.eval()
    if source_img.dim() == 3:
        source_img = source_img.unsqueeze(0)
    if mask.dim() == 3:
        mask = mask.unsqueeze(0)
    if zero_mask:
        mask = torch.zeros_like(mask)
    node = NeuralODE(lambda t, x: net(t, torch.cat([x, mask], dim=1)), solver=solver, sensitivity="adjoint", atol=atol,
                     rtol=rtol)
    t_span = torch.linspace(0, 1, num_steps, device=source_img.device)
    return node.trajectory(source_img, t_span=t_span)[-1]


@torch.no_grad()
def generate(net, source_img, num_steps=100, solver="dopri5", atol=1e-4, rtol=1e-4, y=None):
    """src/models/conditional_flow_matching.py:133-170 (reference default: dopri5, 1e-4)."""
    net.eval()
    if source_img.dim() == 3:
        source_img = source_img.unsqueeze(0)
    vf = net if y is None else (lambda t, x: net(t, x, y=y))
    node = NeuralODE(vf, solver=solver, sensitivity="adjoint", atol=atol, rtol=rtol)
    t_span = torch.linspace(0, 1, num_steps, device=source_img.device)
    return node.trajectory(source_img, t_span=t_span)[-1]


def psnr(a: torch.Tensor, b: torch.Tensor, data_range: float = 2.0) -> float:
    """PSNR between tiles in [-1, 1] (range 2.0)."""
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    if mse == 0:
        return float("inf")
    import math
    return 10.0 * math.log10(data_range ** 2 / mse)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
