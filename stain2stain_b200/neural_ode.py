"""Drop-in for `torchdyn.core.NeuralODE` as the reference uses it:
    node = NeuralODE(self.net, solver=..., sensitivity=..., atol=..., rtol=...)        conditional_flow_matching.py:157-163
    traj = node.trajectory(x, t_span=torch.linspace(0, 1, num_steps))                   :166-167
`trajectory` returns the state at every `t_span` point, `[len(t_span), *x.shape]`.

Fixed-step Euler over one of this package's UNets takes the B200 fast path: one velocity evaluation + state update
(`x += dt * v`, fused into the head conv's epilogue) is captured ONCE in a CUDA graph and replayed per step; time lives
in device memory so the graph is step-invariant and there is no host synchronisation inside the loop.
Every other (solver, vector field) combination runs the generic stepping code below on the caller's vector field.
"""
from __future__ import annotations

import weakref
from typing import Callable, Dict, Optional, Tuple

import torch
import torch.nn as nn

from .unet import RawUNetModel

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
_DOPRI_B5 = _DOPRI_A[6] + (0.0,)
_DOPRI_B4 = (5179 / 57600, 0.0, 7571 / 16695, 393 / 640, -92097 / 339200, 187 / 2100, 1 / 40)


def _rms(x):
    return x.abs().pow(2).mean().sqrt()


def _unwrap_vector_field(vf) -> Tuple[Optional[RawUNetModel], Optional[torch.Tensor], Optional[torch.Tensor]]:
    """Find one of our UNets behind the reference's wrappers -- `ConditionalWrapper(model, y)`
    (class_conditional_flow_matching.py:163-174), `FlowWrapper`, `MaskConditionedWrapper(net, mask)`
    (conditional_flow_matching_conditional_mask.py:170-180) -- returning (net, y, mask)."""
    if isinstance(vf, RawUNetModel):
        return vf, None, None
    inner = getattr(vf, "model", None)
    if inner is None:
        inner = getattr(vf, "net", None)
    if isinstance(inner, RawUNetModel):
        y = getattr(vf, "y", None)
        mask = getattr(vf, "mask", None)
        return inner, (y if torch.is_tensor(y) else None), (mask if torch.is_tensor(mask) else None)
    return None, None, None


class _EulerGraph:
    """One captured Euler step of a UNet on static buffers."""

    def __init__(self, net: RawUNetModel, x: torch.Tensor, y: Optional[torch.Tensor], dt: float,
                 extra: Optional[torch.Tensor] = None):
        from . import kernels as K
        self.x = torch.empty_like(x)
        self.t = torch.zeros(x.shape[0], dtype=torch.float32, device=x.device)
        self.y = None if y is None else y.clone()
        self.extra = None if extra is None else extra.float().contiguous().clone()
        self.dt = dt
        # eager warm-up on scratch data: packs weights, sets kernel attributes, warms the allocator
        self.x.copy_(x)
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            net.euler_step_(self.t, self.x, dt, y=self.y, extra=self.extra)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        n0 = K.LAUNCHES[0]
        with torch.cuda.graph(self.graph):
            net.euler_step_(self.t, self.x, dt, y=self.y, extra=self.extra)
            self.t.add_(dt)
        self.kernels_per_replay = K.LAUNCHES[0] - n0  # this package's kernels inside one replay

    def run(self, x0: torch.Tensor, t0: float, steps: int, y: Optional[torch.Tensor], record: Optional[list],
            extra: Optional[torch.Tensor] = None):
        from . import kernels as K
        self.x.copy_(x0)
        self.t.fill_(t0)
        if self.y is not None:
            self.y.copy_(y)
        if self.extra is not None:
            self.extra.copy_(extra)
        for _ in range(steps):
            self.graph.replay()
            if record is not None:
                record.append(self.x.clone())
        K.LAUNCHES[0] += steps * self.kernels_per_replay
        return self.x


# captured Euler steps per net: a WeakKeyDictionary, so a freed model releases its graphs (and their memory pools) and a
# new model that happens to reuse the old one's id() can never replay another model's capture
_GRAPHS: "weakref.WeakKeyDictionary[nn.Module, Dict[tuple, _EulerGraph]]" = weakref.WeakKeyDictionary()


def clear_graphs():
    """Drop every captured sampler step (call after writing parameters through `.data`, which bumps no version)."""
    _GRAPHS.clear()


def _param_signature(net: nn.Module):
    return tuple((p.data_ptr(), p._version) for p in net.parameters())


def fused_euler(net: RawUNetModel, x: torch.Tensor, t_span: torch.Tensor, y: Optional[torch.Tensor] = None,
                record: Optional[list] = None, use_graph: bool = True, extra: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Integrate dx/dt = net(t, x[, mask]) over a UNIFORM t_span with the fused Euler step.  Returns the final state."""
    assert not net.training, "sampling runs in eval mode (generate() calls self.eval())"
    steps = len(t_span) - 1
    t0 = float(t_span[0])
    dt = float(t_span[-1] - t_span[0]) / steps
    x = x.float().contiguous()
    if use_graph:
        key = (tuple(x.shape), x.device.index, dt, y is not None, extra is not None, _param_signature(net))
        per_net = _GRAPHS.setdefault(net, {})
        g = per_net.get(key)
        if g is None:
            for k in [k for k in per_net if k[0] == tuple(x.shape)]:
                del per_net[k]  # stale parameters / other dt: drop the old capture
            g = per_net[key] = _EulerGraph(net, x, y, dt, extra)
        return g.run(x, t0, steps, y, record, extra).clone()
    xs = x.clone()
    ex = None if extra is None else extra.float().contiguous()
    t = torch.full((x.shape[0],), t0, dtype=torch.float32, device=x.device)
    for _ in range(steps):
        net.euler_step_(t, xs, dt, y=y, extra=ex)
        t += dt
        if record is not None:
            record.append(xs.clone())
    return xs


def _is_uniform(t_span: torch.Tensor) -> bool:
    if len(t_span) < 2:
        return False
    d = t_span[1:] - t_span[:-1]
    return bool((d - d[0]).abs().max() <= 1e-6 * max(1.0, float(d[0].abs())))


def odeint(f: Callable, x: torch.Tensor, t_span: torch.Tensor, solver: str = "dopri5", atol: float = 1e-4,
           rtol: float = 1e-4):
    """Generic stepping on an arbitrary vector field `f(t, x)` (0-dim tensor t).  Returns (t_span, states)."""
    t_span = t_span.to(x.device, torch.float32)
    sol = [x]
    if solver in ("euler", "midpoint", "rk4"):
        for k in range(len(t_span) - 1):
            t, dt = t_span[k], t_span[k + 1] - t_span[k]
            if solver == "euler":
                x = x + dt * f(t, x)
            elif solver == "midpoint":
                x = x + dt * f(t + 0.5 * dt, x + 0.5 * dt * f(t, x))
            else:
                k1 = f(t, x)
                k2 = f(t + dt / 3, x + dt * k1 / 3)
                k3 = f(t + dt * 2 / 3, x + dt * (k2 - k1 / 3))
                k4 = f(t + dt, x + dt * (k1 - k2 + k3))
                x = x + dt * (k1 + 3 * (k2 + k3) + k4) / 8
            sol.append(x)
        return t_span, torch.stack(sol)
    if solver not in ("dopri5",):
        raise NotImplementedError(f"solver {solver!r}: available euler, midpoint, rk4, dopri5")
    # torchdyn's adaptive stepping (B.2): init_step with exponent 1/(order+1), order 5; steps cut at every t_span point;
    # after a cut the controller resumes from what was left of the un-cut proposal; accepted steps never shrink.
    order, safety, fmin, fmax = 5, 0.9, 0.2, 10.0
    t = t_span[0]
    k1 = f(t, x)
    scale = atol + x.abs() * rtol
    d0, d1 = _rms(x / scale), _rms(k1 / scale)
    h0 = x.new_tensor(1e-6) if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    d2 = _rms((f(t + h0, x + h0 * k1) - k1) / scale) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = torch.max(x.new_tensor(1e-6), h0 * 1e-3)
    else:
        h1 = x.new_tensor((0.01 / max(float(d1), float(d2))) ** (1.0 / (order + 1)))
    dt = torch.min(100 * h0, h1)
    nxt = 1
    while nxt < len(t_span):
        stop = t_span[nxt]
        lands = bool(t + dt > stop)
        h = (stop - t) if lands else dt
        ks = [k1]
        for s in range(1, 7):
            ks.append(f(t + _DOPRI_C[s] * h, x + h * sum(a * k for a, k in zip(_DOPRI_A[s], ks))))
        x_new = x + h * sum(b * k for b, k in zip(_DOPRI_B5, ks))
        err = h * sum((b5 - b4) * k for b5, b4, k in zip(_DOPRI_B5, _DOPRI_B4, ks))
        ratio = float(_rms(err / (atol + rtol * torch.max(x.abs(), x_new.abs()))))
        if ratio <= 1.0:
            t, x, k1 = (stop if lands else t + h), x_new, ks[-1]
            if lands:
                sol.append(x)
                nxt += 1
        rest = (dt - h) if lands else h  # a cut step continues from the remainder of its proposal
        if float(rest) <= 0.0:  # (guard: a remainder rounded to 0 would stall the loop)
            rest = h
        if ratio == 0.0:
            grow = fmax
        else:
            grow = min(fmax, max(1.0 if ratio < 1.0 else fmin, safety * ratio ** (-1.0 / order)))
        dt = rest * grow
    return t_span, torch.stack(sol)


class NeuralODE(nn.Module):
    """`torchdyn.core.NeuralODE(vector_field, solver, order, atol, rtol, sensitivity, ...)`; only what `generate`
    touches is implemented (no adjoint backward: the reference never differentiates through `trajectory`)."""

    def __init__(self, vector_field, solver="tsit5", order=1, atol=1e-3, rtol=1e-3, sensitivity="autograd",
                 solver_adjoint=None, atol_adjoint=1e-4, rtol_adjoint=1e-4, interpolator=None, integral_loss=None,
                 seminorm=False, return_t_eval=True, optimizable_params=(), use_cuda_graph=True):
        super().__init__()
        self.vf = vector_field
        self.solver, self.order, self.atol, self.rtol, self.sensitivity = solver, order, atol, rtol, sensitivity
        self.return_t_eval = return_t_eval
        self.use_cuda_graph = use_cuda_graph

    def _fast(self, t_span, x=None):
        net, y, mask = _unwrap_vector_field(self.vf)
        ok = net is not None and self.solver == "euler" and not net.training and _is_uniform(t_span)
        if ok and y is not None and x is not None and tuple(y.shape) != (x.shape[0],):
            ok = False  # a wrapper that broadcasts / truncates its labels per call: let it (generic path)
        return (net, y, mask) if ok else (None, None, None)

    @torch.no_grad()
    def final_state(self, x: torch.Tensor, t_span: torch.Tensor) -> torch.Tensor:
        """State at t_span[-1] without materialising the trajectory (what `generate` actually needs)."""
        net, y, mask = self._fast(t_span, x)
        if net is not None:
            return fused_euler(net, x, t_span, y, None, self.use_cuda_graph, extra=mask)
        return self.trajectory(x, t_span)[-1]

    def trajectory(self, x: torch.Tensor, t_span: torch.Tensor) -> torch.Tensor:
        net, y, mask = self._fast(t_span, x)
        if net is not None and not torch.is_grad_enabled():
            rec = [x.float().clone()]
            fused_euler(net, x, t_span, y, rec, self.use_cuda_graph, extra=mask)
            return torch.stack(rec)
        _, sol = odeint(lambda t, z: self.vf(t, z), x, t_span, solver=self.solver, atol=self.atol, rtol=self.rtol)
        return sol

    def forward(self, x: torch.Tensor, t_span: torch.Tensor):
        sol = self.trajectory(x, t_span)
        return (t_span, sol) if self.return_t_eval else sol
