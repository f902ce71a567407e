"""`FusedAdam`: drop-in for the `torch.optim.Adam` the reference configures as a Hydra `_partial_`
(configs/model/conditional_flow_matching.yaml:3-7; instantiated at src/models/conditional_flow_matching.py:117 as
`self.optimizer(params=self.parameters())`).

Same constructor keywords, same update rule (L2 weight-decay form, bias correction, eps outside the sqrt), same
`state_dict()` layout (`state[p] = {step, exp_avg, exp_avg_sq}`, so Lightning checkpoints interoperate), but the whole
model is updated by ONE launch of the multi-tensor kernel (csrc/optim.cuh) instead of torch's foreach kernel chain.
CUDA fp32 parameters only; anything else raises (no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch

from . import _lib
from . import kernels as K


class _AdamTensor(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("n", C.c_longlong)]


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 amsgrad: bool = False, grad_scale: float = 1.0, **unused):
        if amsgrad:
            raise NotImplementedError("amsgrad is not used by the reference configs")
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self.grad_scale = float(grad_scale)
        self._tables = {}  # group index -> (signature, device tensor table, device work list, n_work, keep-alive)

    def _table(self, gi, ps: List[torch.Tensor], gs: Optional[List[torch.Tensor]] = None):
        """Device pointer table of one group; `gs` = the gradient tensors (default: each parameter's `.grad`)."""
        if gs is None:
            gs = [p.grad for p in ps]
        # every pointer the device table holds is part of the signature: `load_state_dict` / a state reset swaps the moment
        # tensors while parameters and gradients stay where they are
        sig = tuple((p.data_ptr(), g.data_ptr(), self.state[p]["exp_avg"].data_ptr(),
                     self.state[p]["exp_avg_sq"].data_ptr()) for p, g in zip(ps, gs))
        hit = self._tables.get(gi)
        if hit is not None and hit[0] == sig:
            return hit
        dev = ps[0].device
        chunk = _lib.load().s2s_adam_chunk()
        arr = (_AdamTensor * len(ps))()
        work = []
        for i, (p, g) in enumerate(zip(ps, gs)):
            st = self.state[p]
            arr[i] = _AdamTensor(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                 p.numel())
            work.extend((i, c) for c in range((p.numel() + chunk - 1) // chunk))
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone().pin_memory()
        t_dev = raw.to(dev, non_blocking=True)
        w_host = torch.tensor(work, dtype=torch.int32).pin_memory()
        w_dev = w_host.to(dev, non_blocking=True)
        hit = (sig, t_dev, w_dev, len(work), (raw, w_host))
        self._tables[gi] = hit
        return hit

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._tables.clear()
        for st in self.state.values():  # torch casts `step` to the parameter's device; the host-side count stays on the CPU
            if torch.is_tensor(st.get("step")) and st["step"].is_cuda:
                st["step"] = st["step"].detach().cpu()

    def add_param_group(self, param_group):
        super().add_param_group(param_group)
        if hasattr(self, "_tables"):
            self._tables.clear()

    def _init_state(self, p):
        st = self.state[p]
        if not st:
            st["step"] = torch.zeros((), dtype=torch.float32)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    @torch.no_grad()
    def step_on_device_counter(self, step_dev: torch.Tensor, grads: Optional[Dict[torch.Tensor, torch.Tensor]] = None,
                               grad_scale: Optional[float] = None):
        """The same update with the step count read from DEVICE memory (`step_dev`: int64 scalar, already incremented),
        for launches captured in a CUDA graph (graphed.py): nothing host-side changes between replays, so the bias
        corrections are evaluated on the device.  `grads`: parameter -> gradient tensor to read instead of `.grad` (views
        of the flat all-reduced buffer).  The host-side `state[p]["step"]` is NOT advanced here; the caller accounts for it
        (`advance_steps`)."""
        assert step_dev.is_cuda and step_dev.dtype == torch.int64 and step_dev.numel() == 1
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if (p in grads if grads is not None else p.grad is not None)]
            if not ps:
                continue
            gs = [grads[p] if grads is not None else p.grad for p in ps]
            for p, g in zip(ps, gs):
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise _lib.S2SError("FusedAdam needs contiguous fp32 CUDA parameters (no CPU fallback)")
                if g.dtype != torch.float32 or not g.is_contiguous() or g.numel() != p.numel():
                    raise _lib.S2SError("FusedAdam needs dense contiguous fp32 gradients")
                self._init_state(p)
            _, t_dev, w_dev, n_work, _ = self._table(("dev", gi, grads is not None), ps, gs)
            b1, b2 = group["betas"]
            with K._Prof("adam_multi", 0.0, 28.0 * sum(p.numel() for p in ps)):
                K.check(_lib.load().s2s_adam_multi_step(t_dev.data_ptr(), w_dev.data_ptr(), n_work, float(group["lr"]),
                                                        float(b1), float(b2), float(group["eps"]),
                                                        float(group["weight_decay"]), 1, step_dev.data_ptr(),
                                                        self.grad_scale if grad_scale is None else float(grad_scale),
                                                        _lib.stream_ptr()), "adam_multi_step")
            torch.autograd.graph.increment_version(ps)

    def advance_steps(self, n: int):
        """Account on the host for `n` updates that ran from a CUDA graph (keeps `state_dict()` = torch.optim.Adam's)."""
        for st in self.state.values():
            if "step" in st:
                st["step"] += n

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise _lib.S2SError("FusedAdam needs contiguous fp32 CUDA parameters (no CPU fallback)")
                g = p.grad
                if g.is_sparse or g.dtype != torch.float32 or not g.is_contiguous():
                    raise _lib.S2SError("FusedAdam needs dense contiguous fp32 gradients")
                self._init_state(p)
            steps = {int(self.state[p]["step"]) for p in ps}
            if len(steps) != 1:
                raise _lib.S2SError("FusedAdam: parameters of one group must share a step count")
            step = steps.pop() + 1
            _, t_dev, w_dev, n_work, _ = self._table(gi, ps)
            b1, b2 = group["betas"]
            with K._Prof("adam_multi", 0.0, 28.0 * sum(p.numel() for p in ps)):
                K.check(_lib.load().s2s_adam_multi(t_dev.data_ptr(), w_dev.data_ptr(), n_work, float(group["lr"]),
                                                   float(b1), float(b2), float(group["eps"]),
                                                   float(group["weight_decay"]), step, self.grad_scale,
                                                   _lib.stream_ptr()), "adam_multi")
            # the kernel wrote through raw pointers: tell autograd (and the packed-operand cache, which is keyed by
            # parameter version) that the parameters changed
            torch.autograd.graph.increment_version(ps)
            for p in ps:
                self.state[p]["step"] += 1
        return loss
