"""B200 drop-ins for the reference's multitask model (BASELINE config 5 / SURVEY.md rows a18-a20):

  SharedEncoder, DoubleConv, Down, TimeEmbedding ........ src/models/components/shared_encoder.py:40-104, 9-24, 27-37, 107-135
  Up, FlowMatchingDecoder, SegmentationDecoder .......... src/models/components/task_decoders.py:27-50, 53-134, 137-194
  MultiTaskFlowMatchingLitModule ......................... src/models/conditional_flow_matching_multitask_multiclassloss.py:86-573
  (yaml: configs/model/conditional_flow_matching_multitask_multiclass.yaml)

Same class names, constructor kwargs, module tree and therefore `state_dict()` keys (incl. BatchNorm running
statistics) as the reference.  The nn.Conv2d / nn.BatchNorm2d children are parameter containers; the math runs in the
sm_100a kernels: tcgen05 implicit-GEMM convs (the decoder's `cat([skip, up])` is two GEMM segments, never
materialised), train-mode BatchNorm + ReLU as two streaming passes with a batch-wide fold, max-pool / bilinear x2
(align_corners) kernels, one fused softmax-Dice + cross-entropy loss.  Tensors between encoder and decoders are 16-bit
NHWC; images, velocities and logits are fp32 NCHW exactly as the reference passes them.  CUDA only, no CPU fallback.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .flow_matching import ConditionalFlowMatcher
from .lit import _Base, _solver_attr
from .neural_ode import NeuralODE
from .ops import ConvPlan, Seg


def _need_cuda(x):
    if not x.is_cuda:
        raise RuntimeError("stain2stain_b200.multitask runs on CUDA (sm_100a) only; there is no CPU fallback")


class DoubleConv(nn.Module):
    """(conv3x3 => BatchNorm2d => ReLU) * 2; `forward(srcs)` takes the list of NHWC tensors whose channel concat is the
    block input, or a fp32 NCHW image for the 3-channel stem."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.double_conv = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1), nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True),
            nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1), nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True))
        self.in_channels, self.out_channels = in_channels, out_channels
        if out_channels % 64:
            raise NotImplementedError("the B200 conv kernels need feature widths that are multiples of 64")
        self._plan2 = ConvPlan((Seg(0, 0, 0, out_channels, 9, 1),), out_channels)
        self._plans1: Dict[tuple, ConvPlan] = {}

    def _plan1(self, widths):
        key = tuple(widths)
        if key not in self._plans1:
            segs, off = [], 0
            for i, w in enumerate(widths):
                segs.append(Seg(i, 0, off, w, 9, 1))
                off += w
            assert off == self.in_channels, f"DoubleConv expects {self.in_channels} input channels, got {off}"
            self._plans1[key] = ConvPlan(tuple(segs), self.out_channels)
        return self._plans1[key]

    def forward(self, srcs, x1=None, t=None):
        c1, bn1, _, c2, bn2, _ = self.double_conv
        if torch.is_tensor(srcs):  # fp32 NCHW image (stem); optional fused FM interpolation with (x1, t)
            if self.in_channels != 3:
                raise NotImplementedError("image-space inputs are supported for the 3-channel stem only")
            h = ops.stem_conv(srcs.float().contiguous(), c1.weight, c1.bias, x1=x1, t=t)
        else:
            h = ops.fused_conv(self._plan1([s.shape[3] for s in srcs]), list(srcs), [c1.weight], [c1.bias])
        a = ops.batch_norm_relu(h, bn1)
        h = ops.fused_conv(self._plan2, [a], [c2.weight], [c2.bias])
        return ops.batch_norm_relu(h, bn2)


class Down(nn.Module):
    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), DoubleConv(in_channels, out_channels))

    def forward(self, x):
        return self.maxpool_conv[1]([ops.maxpool2x(x)])


class SharedEncoder(nn.Module):
    def __init__(self, in_channels: int = 3, features: Optional[List[int]] = None, return_skip_connections: bool = True):
        super().__init__()
        features = [64, 128, 256, 512, 1024] if features is None else list(features)
        self.in_channels, self.features, self.return_skip_connections = in_channels, features, return_skip_connections
        self.inc = DoubleConv(in_channels, features[0])
        self.downs = nn.ModuleList(Down(features[i], features[i + 1]) for i in range(len(features) - 1))

    def forward(self, x, x1=None, t=None) -> Tuple[torch.Tensor, List[torch.Tensor]]:
        """x: fp32 NCHW image.  With (x1, t): encodes the interpolant (1-t) x + t x1 without materialising it."""
        _need_cuda(x)
        skips = []
        h = self.inc(x, x1=x1, t=t)
        skips.append(h)
        for down in self.downs:
            h = down(h)
            skips.append(h)
        bottleneck = skips[-1]
        return (bottleneck, skips[:-1][::-1]) if self.return_skip_connections else (bottleneck, [])


class TimeEmbedding(nn.Module):
    """sin first, then cos; exponent denominator (half_dim - 1) (shared_encoder.py:107-135)."""

    def __init__(self, dim: int):
        super().__init__()
        self.dim = dim

    def forward(self, t):
        half = self.dim // 2
        e = torch.log(torch.tensor(10000.0)) / (half - 1)
        e = torch.exp(torch.arange(half, device=t.device) * -e)
        if t.dim() == 1:
            t = t.unsqueeze(-1)
        e = t * e.unsqueeze(0)
        return torch.cat([torch.sin(e), torch.cos(e)], dim=-1)


class Up(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, bilinear: bool = True):
        super().__init__()
        if not bilinear:
            raise NotImplementedError("bilinear=False (ConvTranspose2d) is not used by the reference configs")
        self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
        self.conv = DoubleConv(in_channels, out_channels)

    def forward(self, x1, x2):
        u = ops.bilinear2x(x1)
        if u.shape[1:3] != x2.shape[1:3]:
            raise NotImplementedError("odd feature-map sizes (F.pad path of the reference's Up) are not supported")
        return self.conv([x2, u])  # == cat([x2, x1], dim=1) as two GEMM segments


class _Decoder(nn.Module):
    def _build(self, bottleneck_channels, features, out_channels, bilinear):
        features = [512, 256, 128, 64] if features is None else list(features)
        self.ups = nn.ModuleList()
        in_ch = bottleneck_channels
        for f in features:
            self.ups.append(Up(in_ch + f, f, bilinear))
            in_ch = f
        self.outc = nn.Conv2d(features[-1], out_channels, kernel_size=1)

    def _run(self, x, skips):
        for up, skip in zip(self.ups, skips):
            x = up(x, skip)
        return ops.head_conv1x1(x, self.outc.weight, self.outc.bias)


class FlowMatchingDecoder(_Decoder):
    def __init__(self, bottleneck_channels: int = 1024, features: Optional[List[int]] = None, out_channels: int = 3,
                 time_emb_dim: int = 256, bilinear: bool = True):
        super().__init__()
        self.bottleneck_channels, self.time_emb_dim = bottleneck_channels, time_emb_dim
        self.time_mlp = nn.Sequential(nn.Linear(time_emb_dim, time_emb_dim), nn.SiLU(),
                                      nn.Linear(time_emb_dim, time_emb_dim))
        self.time_proj = nn.Linear(time_emb_dim, bottleneck_channels)
        self._build(bottleneck_channels, features, out_channels, bilinear)

    def forward(self, bottleneck, skip_connections, t_emb):
        t = self.time_proj(self.time_mlp(t_emb.float()))
        return self._run(ops.channel_bias_add(bottleneck, t), skip_connections)


class SegmentationDecoder(_Decoder):
    def __init__(self, bottleneck_channels: int = 1024, features: Optional[List[int]] = None, out_channels: int = 1,
                 bilinear: bool = True):
        super().__init__()
        self._build(bottleneck_channels, features, out_channels, bilinear)

    def forward(self, bottleneck, skip_connections):
        return self._run(bottleneck, skip_connections)


class MulticlassDiceLoss(nn.Module):
    """Same constructor as the reference's class (:31-38); generic torch fallback for non-CUDA callers is NOT provided:
    the fused kernel path is `ops.seg_loss`."""

    def __init__(self, num_classes: int, smooth: float = 1.0, ignore_index: int = -100):
        super().__init__()
        self.num_classes, self.smooth, self.ignore_index = num_classes, smooth, ignore_index

    def forward(self, pred, target):
        _need_cuda(pred)
        _, dice, _ = ops.seg_loss(pred.float(), target.long(), self.num_classes, self.ignore_index, 1.0, self.smooth)
        return dice


class FlowWrapper(nn.Module):
    """generate()'s inner wrapper (:541-554): broadcasts t and routes to forward_flow."""

    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, t, x, **kwargs):
        if t.dim() == 0:
            t = t.unsqueeze(0).expand(x.shape[0])
        elif t.dim() == 1 and t.shape[0] == 1:
            t = t.expand(x.shape[0])
        return self.model.forward_flow(t, x)


class MultiTaskFlowMatchingLitModule(_Base):
    def __init__(self, encoder: SharedEncoder, flow_decoder: FlowMatchingDecoder, seg_decoder: SegmentationDecoder,
                 flow_matcher: ConditionalFlowMatcher, num_classes: int = 2, solver: Optional[Any] = None,
                 optimizer: Any = None, scheduler: Any = None, compile: bool = False, log_images: bool = True,
                 seg_loss_weight: float = 1.0, dice_weight: float = 0.5, n_images_log: int = 5, time_emb_dim: int = 256,
                 ignore_index: int = -100, reference_solver_defaults: bool = False) -> None:
        super().__init__()
        self.save_hyperparameters(logger=False)
        self.encoder, self.flow_decoder, self.seg_decoder = encoder, flow_decoder, seg_decoder
        self.time_embedding = TimeEmbedding(time_emb_dim)
        self.flow_matcher, self.solver = flow_matcher, solver
        self.optimizer, self.scheduler = optimizer, scheduler
        self.num_classes, self.ignore_index = num_classes, ignore_index
        self.dice_loss = MulticlassDiceLoss(num_classes=num_classes, ignore_index=ignore_index)
        self.ce_loss = nn.CrossEntropyLoss(ignore_index=ignore_index)
        self.seg_loss_weight, self.dice_weight = seg_loss_weight, dice_weight
        self.log_images, self.n_images_log = log_images, n_images_log
        self.reference_solver_defaults = reference_solver_defaults

    # ------------------------------------------------------------------ forward passes
    def forward_flow(self, t: torch.Tensor, x: torch.Tensor, x1: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Velocity at (t, x).  With `x1` the encoder input is the interpolant (1-t) x + t x1 (fused in the stem)."""
        tt = t
        while tt.dim() > 1:
            tt = tt[:, 0]
        tt = tt.to(x.device).float()
        bottleneck, skips = self.encoder(x, x1=None if x1 is None else x1.float().contiguous(),
                                         t=None if x1 is None else tt.contiguous())
        return self.flow_decoder(bottleneck, skips, self.time_embedding(t.to(x.device)))

    def forward_segmentation(self, x: torch.Tensor) -> torch.Tensor:
        bottleneck, skips = self.encoder(x)
        return self.seg_decoder(bottleneck, skips)

    def compute_segmentation_loss(self, pred_mask, target_mask):
        if target_mask.dim() == 4 and target_mask.shape[1] == 1:
            target_mask = target_mask.squeeze(1)
        target_mask = target_mask.long()
        seg, dice, ce = ops.seg_loss(pred_mask.float(), target_mask, self.num_classes, self.ignore_index,
                                     self.dice_weight, self.dice_loss.smooth)
        return seg, {"dice": dice, "ce": ce, "seg_total": seg}

    def model_step(self, batch, t: Optional[torch.Tensor] = None):
        source_img, target_img, gt_mask = batch
        x0, x1 = source_img, target_img
        if float(getattr(self.flow_matcher, "sigma", 1.0)) == 0.0 and hasattr(self.flow_matcher, "sample_time"):
            if t is None:
                t = self.flow_matcher.sample_time(x0)
            vt = self.forward_flow(t, x0, x1=x1)
            flow_loss = ops.fm_loss(vt.float(), x0.float(), x1.float())
        else:
            t, xt, ut = self.flow_matcher.sample_location_and_conditional_flow(x0, x1, t=t)
            vt = self.forward_flow(t, xt)
            flow_loss = torch.mean((vt - ut) ** 2)
        pred_mask_logits = self.forward_segmentation(source_img)  # second pass through the shared encoder
        seg_loss, d = self.compute_segmentation_loss(pred_mask_logits, gt_mask)
        total = flow_loss + self.seg_loss_weight * seg_loss
        return total, {"total": total, "flow": flow_loss, "seg": seg_loss, "seg_dice": d["dice"], "seg_ce": d["ce"]}

    def _log_all(self, stage, d, on_step):
        for key, name, bar in (("total", "loss", True), ("flow", "flow_loss", False), ("seg", "seg_loss", True),
                               ("seg_dice", "seg_dice", False), ("seg_ce", "seg_ce", False)):
            self.log(f"{stage}/{name}", d[key], on_step=on_step, on_epoch=True, prog_bar=bar, sync_dist=True)

    def training_step(self, batch, batch_idx: int) -> torch.Tensor:
        total, d = self.model_step(batch)
        self._log_all("train", d, True)
        return total

    def validation_step(self, batch, batch_idx: int) -> None:
        _, d = self.model_step(batch)
        self._log_all("val", d, False)

    def test_step(self, batch, batch_idx: int) -> None:
        _, d = self.model_step(batch)
        self._log_all("test", d, False)

    def configure_optimizers(self) -> Dict[str, Any]:
        optimizer = self.optimizer(params=self.parameters())
        if self.scheduler is not None:
            scheduler = self.scheduler(optimizer=optimizer)
            return {"optimizer": optimizer,
                    "lr_scheduler": {"scheduler": scheduler, "monitor": "val/loss", "interval": "epoch", "frequency": 1}}
        return {"optimizer": optimizer}

    @torch.no_grad()
    def generate(self, source_img: torch.Tensor, num_steps: int = 100) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.solver is None:
            raise ValueError("Solver is not initialized. Cannot perform inference.")
        self.eval()
        if source_img.dim() == 3:
            source_img = source_img.unsqueeze(0)
        device = source_img.device
        pred_mask_probs = F.softmax(self.forward_segmentation(source_img), dim=1)
        pred_mask = torch.argmax(pred_mask_probs, dim=1, keepdim=True)
        if self.reference_solver_defaults:
            get = lambda n, d: getattr(self.solver, n) if hasattr(self.solver, n) else d  # noqa: E731
        else:
            get = lambda n, d: _solver_attr(self.solver, n, d)  # noqa: E731
        node = NeuralODE(FlowWrapper(self), solver=get("solver", "dopri5"), sensitivity=get("sensitivity", "adjoint"),
                         atol=get("atol", 1e-4), rtol=get("rtol", 1e-4))
        t_span = torch.linspace(0, 1, num_steps, device=device)
        traj = node.trajectory(source_img, t_span=t_span)
        return traj[-1], pred_mask

    def on_train_epoch_end(self) -> None:
        return None

    def on_validation_epoch_end(self) -> None:
        return None
