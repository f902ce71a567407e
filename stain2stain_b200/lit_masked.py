"""Mask / ROI variants of the flow-matching LitModule.  In the reference each lives in its own
file under the SAME class name `ConditionalFlowMatchingLitModule`; here they get distinct names and
`hydra_lite.TARGET_MAP` sends each reference `_target_` to its mirror:

  MaskWeightedFlowMatchingLitModule      <- src/models/conditional_flow_matching_masked.py:9-316
      3-channel UNet; loss = sum(w (v-u)^2) / (sum(w) + 1e-8), w = 1 + roi_lambda * mask (default 10, read with
      getattr like the reference :78).
  ROILossFlowMatchingLitModule           <- src/models/conditional_flow_matching_ROI_loss.py:9-362
      3-channel UNet; loss = MSE + lambda_roi * ROI-Charbonnier(xt, x1; mask) (:73-97; the second term carries no
      gradient to the network -- it compares the interpolant with the target).
  MaskConditionedFlowMatchingLitModule   <- src/models/conditional_flow_matching_conditional_mask.py:14-366
      4-channel raw UNet (configs/model/conditional_flow_matching_masked_condition.yaml:18-30): `torch.cat([x, mask], 1)`
      (:54-66) is never materialised -- the mask enters the stem operand packing as the 4th patch channel, in training
      (fused with the interpolation) and inside the CUDA-graph Euler sampler.
  MaskToggleFlowMatchingLitModule        <- src/models/conditional_flow_matching_conditional_toggle_mask.py:14-385
      the same with the mask zeroed with probability 0.5 per training step (:77-79, host RNG `torch.rand(1).item()`
      like the reference) and ALWAYS zeroed in `generate` (:186-187).

Constructor keywords, method names, error behaviour and `state_dict` keys (`net.*`) are the reference's.
"""
from __future__ import annotations

from typing import Any, Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from .lit import ConditionalFlowMatchingLitModule


class _MaskBase(ConditionalFlowMatchingLitModule):
    def __init__(self, net: torch.nn.Module, flow_matcher, solver: Optional[Any] = None, optimizer: Any = None,
                 scheduler: Any = None, compile: bool = False, log_images: bool = True, aux_loss_weight: float = 0.1,
                 n_images_log: int = 5, reference_solver_defaults: bool = False) -> None:
        super().__init__(net, flow_matcher, solver, optimizer, scheduler, compile, log_images, n_images_log,
                         reference_solver_defaults)
        self.aux_loss_weight = aux_loss_weight

    def _sample_t(self, x0, t):
        return self.flow_matcher.sample_time(x0) if t is None else t


class MaskWeightedFlowMatchingLitModule(_MaskBase):
    def __init__(self, net, flow_matcher, solver=None, optimizer=None, scheduler=None, compile: bool = False,
                 log_images: bool = True, aux_loss_weight: float = 0.1, reference_solver_defaults: bool = False) -> None:
        super().__init__(net, flow_matcher, solver, optimizer, scheduler, compile, log_images, aux_loss_weight, 5,
                         reference_solver_defaults)

    def model_step(self, batch: Tuple[torch.Tensor, ...], t: Optional[torch.Tensor] = None) -> torch.Tensor:
        x0, x1, mask = batch
        lam = getattr(self, "roi_lambda", 10.0)
        if self._fused_ok() and x0.is_cuda:
            t = self._sample_t(x0, t)
            vt = self.net.velocity_of_interpolant(t, x0, x1)
            return ops.fm_loss_weighted(vt.float(), x0.float(), x1.float(), mask.float(), lam)
        t, xt, ut = self.flow_matcher.sample_location_and_conditional_flow(x0, x1, t=t)
        vt = self.forward(t, xt)
        weights = (1.0 + lam * mask).expand_as(vt)
        return (weights * (vt - ut) ** 2).sum() / (weights.sum() + 1e-8)


class ROILossFlowMatchingLitModule(_MaskBase):
    def model_step(self, batch: Tuple[torch.Tensor, ...], t: Optional[torch.Tensor] = None) -> torch.Tensor:
        x0, x1, mask = batch
        lambda_roi = getattr(self, "lambda_roi", 1.0)
        if self._fused_ok() and x0.is_cuda:
            t = self._sample_t(x0, t)
            vt = self.net.velocity_of_interpolant(t, x0, x1)
            loss_fm = ops.fm_loss(vt.float(), x0.float(), x1.float())
            return loss_fm + lambda_roi * ops.roi_charbonnier(x0.float(), x1.float(), t.to(x0.device), mask)
        t, xt, ut = self.flow_matcher.sample_location_and_conditional_flow(x0, x1, t=t)
        vt = self.forward(t, xt)
        loss_fm = torch.mean((vt - ut) ** 2)
        m = mask.float()
        diff = xt - x1
        charb = torch.sqrt(diff * diff + 1e-3 * 1e-3)
        return loss_fm + lambda_roi * (charb * m).sum() / (m.sum() * xt.shape[1] + 1e-8)


class MaskConditionedWrapper(nn.Module):
    """conditional_flow_matching_conditional_mask.py:170-180: closes over the mask for the ODE solver.  The sampler
    recognises `.net` / `.mask` and feeds the mask to the stem kernel instead of concatenating it every evaluation."""

    def __init__(self, net, mask):
        super().__init__()
        self.net = net
        self.mask = mask

    def forward(self, t, x, **kwargs):
        return self.net(t, torch.cat([x, self.mask], dim=1))


class MaskConditionedFlowMatchingLitModule(_MaskBase):
    toggle_in_training = False
    zero_mask_in_generate = False

    def forward(self, t: torch.Tensor, x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:  # type: ignore[override]
        if hasattr(self.net, "_run") and x.is_cuda:
            return self.net._run(t, x, None, extra=mask.float())
        return self.net(t, torch.cat([x, mask], dim=1))

    def model_step(self, batch: Tuple[torch.Tensor, ...], use_mask_toggle: bool = False,
                   t: Optional[torch.Tensor] = None) -> torch.Tensor:
        x0, x1, mask = batch
        if use_mask_toggle and torch.rand(1).item() < 0.5:
            mask = torch.zeros_like(mask)
        if self._fused_ok() and x0.is_cuda:
            t = self._sample_t(x0, t)
            vt = self.net.velocity_of_interpolant(t, x0, x1, extra=mask.float())
            return ops.fm_loss(vt.float(), x0.float(), x1.float())
        t, xt, ut = self.flow_matcher.sample_location_and_conditional_flow(x0, x1, t=t)
        vt = self.forward(t, xt, mask)
        return torch.mean((vt - ut) ** 2)

    def training_step(self, batch, batch_idx: int) -> torch.Tensor:
        loss = self.model_step(batch, use_mask_toggle=self.toggle_in_training)
        self.log("train/loss", loss, on_step=True, on_epoch=True, prog_bar=True, sync_dist=True)
        return loss

    @torch.no_grad()
    def generate(self, source_img: torch.Tensor, mask: torch.Tensor, num_steps: int = 100) -> torch.Tensor:  # type: ignore[override]
        if self.solver is None:
            raise ValueError("Solver is not initialized. Cannot perform inference.")
        self.eval()
        if source_img.dim() == 3:
            source_img = source_img.unsqueeze(0)
        if mask.dim() == 3:
            mask = mask.unsqueeze(0)
        if self.zero_mask_in_generate:
            mask = torch.zeros_like(mask)
        node = self._make_node(MaskConditionedWrapper(self.net, mask.float()))
        t_span = torch.linspace(0, 1, num_steps, device=source_img.device)
        return node.final_state(source_img, t_span)


class MaskToggleFlowMatchingLitModule(MaskConditionedFlowMatchingLitModule):
    toggle_in_training = True
    zero_mask_in_generate = True
