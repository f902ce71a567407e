"""ORACLE (test infrastructure only) -- fp32 PyTorch restatement of the reference's multitask model (config M).

Unlike the torchcfm UNet, every piece restated here lives IN the reference repo and is importable in the build
container, so this file is PINNED: `oracle/make_golden.py` runs the reference's own code
(src/models/components/shared_encoder.py, src/models/components/task_decoders.py and, through `oracle/ref_bridge.py`,
src/models/conditional_flow_matching_multitask_multiclassloss.py) and commits its outputs as tests/golden/*.pt;
tests/test_golden.py checks this restatement against them bit-for-bit on CPU.

Follows:
  DoubleConv / Down / SharedEncoder / TimeEmbedding ... src/models/components/shared_encoder.py:9-24, 27-37, 40-104, 107-135
  Up / FlowMatchingDecoder / SegmentationDecoder ...... src/models/components/task_decoders.py:27-50, 53-134, 137-194
  MulticlassDiceLoss .................................. src/models/conditional_flow_matching_multitask_multiclassloss.py:31-83
  forward_flow / forward_segmentation / compute_segmentation_loss / model_step / generate
                                                        same file :172-193, 195-210, 212-244, 246-299, 506-573
The module tree (attribute names, construction order) is the reference's, so a seeded construction draws the same
initial weights and `state_dict()` keys match the reference's checkpoints.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


class DoubleConv(nn.Module):
    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.double_conv = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1), nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True),
            nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1), nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True))

    def forward(self, x):
        return self.double_conv(x)


class Down(nn.Module):
    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), DoubleConv(in_channels, out_channels))

    def forward(self, x):
        return self.maxpool_conv(x)


class SharedEncoder(nn.Module):
    def __init__(self, in_channels: int = 3, features: Optional[List[int]] = None, return_skip_connections: bool = True):
        super().__init__()
        features = [64, 128, 256, 512, 1024] if features is None else features
        self.in_channels, self.features, self.return_skip_connections = in_channels, features, return_skip_connections
        self.inc = DoubleConv(in_channels, features[0])
        self.downs = nn.ModuleList(Down(features[i], features[i + 1]) for i in range(len(features) - 1))

    def forward(self, x) -> Tuple[torch.Tensor, List[torch.Tensor]]:
        skips = []
        x = self.inc(x)
        skips.append(x)
        for down in self.downs:
            x = down(x)
            skips.append(x)
        bottleneck = skips[-1]
        return (bottleneck, skips[:-1][::-1]) if self.return_skip_connections else (bottleneck, [])


class TimeEmbedding(nn.Module):
    """sin first, then cos; exponent denominator (half_dim - 1) -- NOT torchcfm's embedding."""

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
        x1 = self.up(x1)
        dy, dx = x2.size(2) - x1.size(2), x2.size(3) - x1.size(3)
        x1 = F.pad(x1, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
        return self.conv(torch.cat([x2, x1], dim=1))


class FlowMatchingDecoder(nn.Module):
    def __init__(self, bottleneck_channels: int = 1024, features: Optional[List[int]] = None, out_channels: int = 3,
                 time_emb_dim: int = 256, bilinear: bool = True):
        super().__init__()
        features = [512, 256, 128, 64] if features is None else features
        self.bottleneck_channels, self.time_emb_dim = bottleneck_channels, time_emb_dim
        self.time_mlp = nn.Sequential(nn.Linear(time_emb_dim, time_emb_dim), nn.SiLU(),
                                      nn.Linear(time_emb_dim, time_emb_dim))
        self.time_proj = nn.Linear(time_emb_dim, bottleneck_channels)
        self.ups = nn.ModuleList()
        in_ch = bottleneck_channels
        for f in features:
            self.ups.append(Up(in_ch + f, f, bilinear))
            in_ch = f
        self.outc = nn.Conv2d(features[-1], out_channels, kernel_size=1)

    def forward(self, bottleneck, skip_connections, t_emb):
        t = self.time_proj(self.time_mlp(t_emb))
        x = bottleneck + t.view(t.shape[0], t.shape[1], 1, 1)
        for up, skip in zip(self.ups, skip_connections):
            x = up(x, skip)
        return self.outc(x)


class SegmentationDecoder(nn.Module):
    def __init__(self, bottleneck_channels: int = 1024, features: Optional[List[int]] = None, out_channels: int = 1,
                 bilinear: bool = True):
        super().__init__()
        features = [512, 256, 128, 64] if features is None else features
        self.ups = nn.ModuleList()
        in_ch = bottleneck_channels
        for f in features:
            self.ups.append(Up(in_ch + f, f, bilinear))
            in_ch = f
        self.outc = nn.Conv2d(features[-1], out_channels, kernel_size=1)

    def forward(self, bottleneck, skip_connections):
        x = bottleneck
        for up, skip in zip(self.ups, skip_connections):
            x = up(x, skip)
        return self.outc(x)


def multiclass_dice_loss(pred, target, num_classes: int, smooth: float = 1.0, ignore_index: int = -100):
    """MulticlassDiceLoss.forward (:40-83): sums run over the WHOLE batch; with ignore_index < 0 the mask is all ones."""
    p = F.softmax(pred, dim=1)
    onehot = F.one_hot(target.long(), num_classes=num_classes).permute(0, 3, 1, 2).float()
    if ignore_index >= 0:
        valid = (target != ignore_index).float().unsqueeze(1)
    else:
        valid = torch.ones_like(target).unsqueeze(1).float()
    scores = []
    for c in range(num_classes):
        pc, tc = p[:, c:c + 1] * valid, onehot[:, c:c + 1] * valid
        inter = (pc * tc).sum()
        union = pc.sum() + tc.sum()
        scores.append((2.0 * inter + smooth) / (union + smooth))
    return 1 - torch.stack(scores).mean()


class MultiTaskModel(nn.Module):
    """The numerics of MultiTaskFlowMatchingLitModule (no Lightning): same attribute names -> same state_dict keys."""

    def __init__(self, encoder, flow_decoder, seg_decoder, flow_matcher, num_classes: int = 2,
                 seg_loss_weight: float = 1.0, dice_weight: float = 0.5, time_emb_dim: int = 256,
                 ignore_index: int = -100):
        super().__init__()
        self.encoder, self.flow_decoder, self.seg_decoder = encoder, flow_decoder, seg_decoder
        self.time_embedding = TimeEmbedding(time_emb_dim)
        self.flow_matcher = flow_matcher
        self.num_classes, self.ignore_index = num_classes, ignore_index
        self.seg_loss_weight, self.dice_weight = seg_loss_weight, dice_weight

    def forward_flow(self, t, x):
        bottleneck, skips = self.encoder(x)
        return self.flow_decoder(bottleneck, skips, self.time_embedding(t))

    def forward_segmentation(self, x):
        bottleneck, skips = self.encoder(x)
        return self.seg_decoder(bottleneck, skips)

    def compute_segmentation_loss(self, pred_mask, target_mask):
        if target_mask.dim() == 4 and target_mask.shape[1] == 1:
            target_mask = target_mask.squeeze(1)
        target_mask = target_mask.long()
        dice = multiclass_dice_loss(pred_mask, target_mask, self.num_classes, ignore_index=self.ignore_index)
        ce = F.cross_entropy(pred_mask, target_mask, ignore_index=self.ignore_index)
        seg = self.dice_weight * dice + (1 - self.dice_weight) * ce
        return seg, {"dice": dice, "ce": ce, "seg_total": seg}

    def model_step(self, batch, t=None) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
        x0, x1, gt_mask = batch
        t, xt, ut = self.flow_matcher.sample_location_and_conditional_flow(x0, x1, t=t)
        vt = self.forward_flow(t, xt)
        flow_loss = torch.mean((vt - ut) ** 2)
        seg_loss, d = self.compute_segmentation_loss(self.forward_segmentation(x0), gt_mask)  # second encoder pass
        total = flow_loss + self.seg_loss_weight * seg_loss
        return total, {"total": total, "flow": flow_loss, "seg": seg_loss, "seg_dice": d["dice"], "seg_ce": d["ce"]}

    @torch.no_grad()
    def generate(self, source_img, num_steps: int = 100, solver: str = "dopri5", atol: float = 1e-4, rtol: float = 1e-4):
        from .flow import NeuralODE
        self.eval()
        if source_img.dim() == 3:
            source_img = source_img.unsqueeze(0)
        pred_mask = torch.argmax(F.softmax(self.forward_segmentation(source_img), dim=1), dim=1, keepdim=True)

        def vf(t, x):
            if t.dim() == 0:
                t = t.unsqueeze(0).expand(x.shape[0])
            elif t.dim() == 1 and t.shape[0] == 1:
                t = t.expand(x.shape[0])
            return self.forward_flow(t, x)
        node = NeuralODE(vf, solver=solver, sensitivity="adjoint", atol=atol, rtol=rtol)
        traj = node.trajectory(source_img, torch.linspace(0, 1, num_steps, device=source_img.device))
        return traj[-1], pred_mask


def build(features=(64, 128, 256, 512, 1024), num_classes: int = 5, time_emb_dim: int = 256, flow_matcher=None,
          **kw) -> MultiTaskModel:
    """configs/model/conditional_flow_matching_multitask_multiclass.yaml, in Hydra's instantiation order
    (encoder, flow_decoder, seg_decoder) so that a seeded build reproduces the reference's initial weights."""
    from .flow import ConditionalFlowMatcher
    features = list(features)
    dec = features[:-1][::-1]
    enc = SharedEncoder(3, features, True)
    fd = FlowMatchingDecoder(features[-1], dec, 3, time_emb_dim, True)
    sd = SegmentationDecoder(features[-1], dec, num_classes, True)
    return MultiTaskModel(enc, fd, sd, flow_matcher or ConditionalFlowMatcher(0.0), num_classes=num_classes,
                          time_emb_dim=time_emb_dim, **kw)
