"""The training step as CUDA graphs: FM sample -> UNet forward -> MSE -> backward ->
[gradient all-reduce] -> Adam, captured ONCE and replayed per step.

Why: every shape of the step is static and it consists of ~1200 kernel launches, many of them tens of microseconds long
(the 32x32 levels); launched from Python the GPU idles ~12 ms of a 187 ms step waiting for the next launch
(profiles/r02_step_kernel_times.txt).  Replayed from a graph the launches are back to back.

What makes the step replayable (nothing host-side may change between replays):
  * inputs live in static device buffers (`x0`, `x1`, `t`); `t ~ U(0,1)^B` is still drawn per step from the CPU default
    generator exactly as torchcfm does (`torch.rand(B)`), then copied into the static buffer;
  * dropout masks: the Philox seed of every launch is mixed with a device-resident step counter (`s2s_gn_apply_step`);
  * Adam's bias corrections are evaluated on the device from the same counter (`s2s_adam_multi_step`);
  * the packed 16-bit GEMM operands are re-packed by the captured multi-tensor pack launch at the top of every replay.

Data parallel (world size > 1): the step is TWO graphs with one eager NCCL call in between --
  graph A: forward, backward, one multi-tensor copy of the 282 gradients into ONE flat fp32 buffer;
  eager  : `all_reduce(flat, SUM)` -- a single 284 MB collective over NVLink / NVSwitch (< 1 ms), nothing else is running on
           the SMs while it runs, so the persistent tcgen05 kernels are never descheduled by NCCL's CTAs;
  graph B: Adam reading the flat buffer with `grad_scale = 1 / world` (the mean DDP would have produced).
`p.grad` then holds the LOCAL (un-reduced) gradient; the averaged one is `flat_grad` / world.

After every replay the parameters' version counters are bumped, so eager code that runs afterwards (sampling,
validation) re-packs its operands instead of using ones that are one optimizer step old.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib
from . import kernels as K
from .optim import FusedAdam
from .parallel import FlatGradients


class _CopyTensor(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("n", C.c_longlong)]


class GraphedTrainStep:
    def __init__(self, lit, optimizer: FusedAdam, batch_shape, device, process_group=None, warmup: int = 3,
                 label_shape=None, extra=None):
        """lit: one of this package's flow-matching LitModules (model_step(batch, t=...) -> loss or (loss, dict)); optimizer:
        its FusedAdam; batch_shape: [B, C, H, W] of x0 / x1 per rank; process_group: None = single GPU, else the NCCL group to
        average gradients over (`dist.group.WORLD` for plain data parallel).  Further batch members: `label_shape` (int64
        class labels of the class-conditional module) or `extra` = [(shape, dtype), ...] (e.g. the multitask module's mask)."""
        if not isinstance(optimizer, FusedAdam):
            raise TypeError("GraphedTrainStep needs the FusedAdam of this package (device-side step count)")
        self.lit, self.opt, self.pg = lit, optimizer, process_group
        self.world = dist.get_world_size(process_group) if process_group is not None else 1
        self.device = torch.device(device)
        B = batch_shape[0]
        self.x0 = torch.zeros(batch_shape, dtype=torch.float32, device=self.device)
        self.x1 = torch.zeros_like(self.x0)
        self.t = torch.zeros(B, dtype=torch.float32, device=self.device)
        specs = list(extra or []) + ([(tuple(label_shape), torch.int64)] if label_shape is not None else [])
        self.extras = [torch.zeros(tuple(sh), dtype=dt, device=self.device) for sh, dt in specs]
        self.y = self.extras[-1] if label_shape is not None else None
        self._t_host = torch.zeros(B, dtype=torch.float32).pin_memory()
        self._t_copied = None  # event behind the last asynchronous H2D copy out of `_t_host`
        self.params = [p for p in lit.parameters() if p.requires_grad]
        self.step_dev = torch.zeros((), dtype=torch.int64, device=self.device)
        self.flat_grad = None
        self.loss = None
        self.launches_per_replay = 0
        self.replays = 0
        self._graph_a = self._graph_b = None
        self._capture(warmup)

    # ------------------------------------------------------------------------------------------------ capture
    def _batch(self):
        return (self.x0, self.x1, *self.extras)

    def _fwd_bwd(self):
        self.step_dev.add_(1)
        for p in self.params:
            p.grad = None
        loss = self.lit.model_step(self._batch(), t=self.t)
        if isinstance(loss, tuple):  # the multitask module returns (total, parts)
            loss = loss[0]
        loss.backward()
        return loss.detach()

    def _capture(self, warmup: int):
        lit, opt, dev = self.lit, self.opt, self.device
        lit.train()
        # the device counter continues the optimizer's own step count (a resumed optimizer keeps its bias corrections)
        steps = {int(st["step"]) for st in opt.state.values() if "step" in st}
        if len(steps) > 1:
            raise _lib.S2SError("GraphedTrainStep: parameters disagree about the step count")
        step0 = steps.pop() if steps else 0
        self.step_dev.fill_(step0)
        # the warm-up runs REAL steps on scratch data: snapshot parameters and optimizer state, put them back afterwards
        snap_p = [p.detach().clone() for p in self.params]
        snap_s = [{k: v.detach().clone() for k, v in opt.state[p].items() if torch.is_tensor(v)} if p in opt.state and opt.state[p]
                  else None for p in self.params]
        snap_buf = [b.detach().clone() for b in lit.buffers()]
        K.DROPOUT_STEP_DEV = self.step_dev
        try:
            # eager warm-up on a side stream: fills the operand / table caches, sets kernel attributes, sizes the allocator
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):
                    self.t.uniform_(0, 1)
                    self.x0.uniform_(-1, 1)
                    self.x1.uniform_(-1, 1)
                    self._fwd_bwd()
                    self._gather()
                    if self.world > 1:
                        dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.pg)
                    opt.step_on_device_counter(self.step_dev, self._flat_views, 1.0 / self.world)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            # ---- graph A: forward + backward + flat gradient gather (+ Adam when there is no collective in between)
            n0 = K.LAUNCHES[0]
            self._graph_a = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph_a):
                self.loss = self._fwd_bwd()
                self._gather()
                if self.world == 1:
                    opt.step_on_device_counter(self.step_dev, self._flat_views, 1.0)
            if self.world > 1:
                self._graph_b = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph_b, pool=self._graph_a.pool()):
                    opt.step_on_device_counter(self.step_dev, self._flat_views, 1.0 / self.world)
            self.launches_per_replay = K.LAUNCHES[0] - n0
        finally:
            K.DROPOUT_STEP_DEV = None
        # capture itself executes nothing; undo the warm-up steps (same storage: the graphs' pointers stay valid)
        with torch.no_grad():
            for p, w, st in zip(self.params, snap_p, snap_s):
                p.copy_(w)
                cur = opt.state[p]
                for k in ("exp_avg", "exp_avg_sq"):
                    if st is None:
                        cur[k].zero_()
                    else:
                        cur[k].copy_(st[k])
                cur["step"] = torch.tensor(float(step0), dtype=torch.float32) if st is None else st["step"].cpu()
            for b, w in zip(lit.buffers(), snap_buf):
                b.copy_(w)
        self.step_dev.fill_(step0)
        torch.autograd.graph.increment_version(self.params)
        torch.cuda.synchronize(dev)

    # ------------------------------------------------------------------------------------------------ flat gradients
    def _ensure_flat(self):
        if self.flat_grad is not None:
            return
        self._flat = FlatGradients(self.params)  # layout shared with the CPU-testable host logic (parallel.py)
        self.flat_grad, self._flat_views = self._flat.flat, self._flat.views
        # copy table: the work list never changes; the pointer table is rewritten whenever the gradients move (eager
        # warm-up vs. the graph's private pool).  Its host staging buffers are pinned and allocated HERE, outside any
        # capture; inside a capture only CPU writes into them and one captured H2D copy happen.
        chunk = _lib.load().s2s_adam_chunk()
        work = []
        for i, p in enumerate(self.params):
            work.extend((i, c) for c in range((p.numel() + chunk - 1) // chunk))
        self._n_work = len(work)
        self._work_dev = torch.tensor(work, dtype=torch.int32).to(self.device)
        nbytes = C.sizeof(_CopyTensor) * len(self.params)
        self._tab_host = [torch.zeros(nbytes, dtype=torch.uint8).pin_memory() for _ in range(2)]  # [eager, capture]
        self._tab_dev = [torch.zeros(nbytes, dtype=torch.uint8, device=self.device) for _ in range(2)]
        self._tab_sig = [None, None]

    def _gather(self):
        """ONE launch: every p.grad -> its slice of the flat buffer."""
        self._ensure_flat()
        which = 1 if torch.cuda.is_current_stream_capturing() else 0
        sig = tuple(p.grad.data_ptr() for p in self.params)
        if self._tab_sig[which] != sig:
            if which == 0:  # eager: an earlier, still queued H2D copy may read the staging buffer
                torch.cuda.current_stream(self.device).synchronize()
            arr = (_CopyTensor * len(self.params))()
            for i, p in enumerate(self.params):
                g = p.grad
                assert g is not None and g.dtype == torch.float32 and g.is_contiguous(), "dense fp32 gradients expected"
                arr[i] = _CopyTensor(g.data_ptr(), self._flat_views[p].data_ptr(), p.numel())
            C.memmove(self._tab_host[which].data_ptr(), C.addressof(arr), C.sizeof(arr))
            self._tab_dev[which].copy_(self._tab_host[which], non_blocking=True)
            self._tab_sig[which] = sig
        with K._Prof("grad_gather", 0.0, 8.0 * self.flat_grad.numel()):
            K.check(_lib.load().s2s_copy_multi(self._tab_dev[which].data_ptr(), self._work_dev.data_ptr(), self._n_work,
                                               _lib.stream_ptr()), "copy_multi")

    # ------------------------------------------------------------------------------------------------ replay
    def load_inputs(self, x0, x1, t: Optional[torch.Tensor] = None, y=None, extra=None):
        """Copy one batch into the static buffers (host or device sources; pinned host memory copies asynchronously).
        t: explicit times, else `torch.rand(B)` from the CPU default generator as torchcfm draws them."""
        self.x0.copy_(x0, non_blocking=True)
        self.x1.copy_(x1, non_blocking=True)
        if t is None:
            # the host may run a step ahead of the device: do not redraw into the staging buffer while the previous
            # step's copy out of it is still queued (that step would train on the NEXT step's times)
            if self._t_copied is not None:
                self._t_copied.synchronize()
            torch.rand(self._t_host.shape, out=self._t_host)
            self.t.copy_(self._t_host, non_blocking=True)
            if self._t_copied is None:
                self._t_copied = torch.cuda.Event()
            self._t_copied.record(torch.cuda.current_stream(self.device))
        else:
            self.t.copy_(t, non_blocking=True)
        if self.y is not None and y is not None:
            self.y.copy_(y, non_blocking=True)
        for dst, src in zip(self.extras, extra or []):
            dst.copy_(src, non_blocking=True)

    def replay(self) -> torch.Tensor:
        """One training step on whatever the static buffers hold.  Returns the (static) loss tensor."""
        self._graph_a.replay()
        if self.world > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.pg)
            self._graph_b.replay()
        self.replays += 1
        self.opt.advance_steps(1)
        K.LAUNCHES[0] += self.launches_per_replay
        # eager code that runs after this (sampling, validation, a checkpoint's packed operands) must see new versions
        torch.autograd.graph.increment_version(self.params)
        return self.loss

    def __call__(self, x0, x1, t: Optional[torch.Tensor] = None, y=None, extra=None) -> torch.Tensor:
        self.load_inputs(x0, x1, t, y, extra)
        return self.replay()
