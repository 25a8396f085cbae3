"""Synthetic head outputs for the scoring path: per-level classification logits
[B, A*C_out, H, W], box deltas [B, A*4, H, W] and MEH lambda maps [B, A, H, W] in the NCHW layout
the RetinaNet / SSD heads emit (channel = a*C_out + c, prior n = (h*W + w)*A + a), seeded per
global image id so a pool is reproducible for any sharding.

There is no network for datasets or checkpoints, so this is what the pool is made of:
  objects   n_gt ~ U{gt_range}, side U(0.05, 0.6)*min(H, W), uniform position, class U{0..C-1}
  logits    N(0,1) background (SSD: background-class logit +4); priors whose anchor has
            IoU >= 0.4 with an object get logit[class] += U(2, 8)
  deltas    positives: bbox2delta(anchor, gt)/std + N(0, 0.05); others N(0, 0.1)
  lambda    relu(N(0.5, 0.3)) + 1e-3   (strictly > 0 so every Dirichlet alpha is > 0)
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from .anchors import grid_anchors
from .specs import HEAD_SSD, DetectorSpec


def _pairwise_iou(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = torch.maximum(a[:, None, :2], b[None, :, :2])
    rb = torch.minimum(a[:, None, 2:], b[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    return inter / (area_a[:, None] + area_b[None, :] - inter).clamp(min=1e-6)


def _encode(anchors: torch.Tensor, gt: torch.Tensor, stds: Sequence[float]) -> torch.Tensor:
    pw = anchors[:, 2] - anchors[:, 0]
    ph = anchors[:, 3] - anchors[:, 1]
    px = (anchors[:, 0] + anchors[:, 2]) * 0.5
    py = (anchors[:, 1] + anchors[:, 3]) * 0.5
    gw = gt[:, 2] - gt[:, 0]
    gh = gt[:, 3] - gt[:, 1]
    gx = (gt[:, 0] + gt[:, 2]) * 0.5
    gy = (gt[:, 1] + gt[:, 3]) * 0.5
    d = torch.stack([(gx - px) / pw, (gy - py) / ph, torch.log(gw / pw), torch.log(gh / ph)], dim=-1)
    return d / d.new_tensor(list(stds))


def _to_nchw(x_nc: torch.Tensor, fh: int, fw: int, a: int) -> torch.Tensor:
    """[N_s, K] with n = (h*W+w)*A + a  ->  [A*K, H, W] with channel = a*K + k."""
    k = x_nc.shape[1]
    return x_nc.view(fh, fw, a * k).permute(2, 0, 1).contiguous()


class SyntheticPool:
    """Generates head outputs for arbitrary global image ids on one device."""

    def __init__(self, spec: DetectorSpec, seed0: int = 20, device="cpu",
                 scale_factor: Sequence[float] = (1.0, 1.0, 1.0, 1.0)):
        self.spec = spec
        self.seed0 = int(seed0)
        self.device = torch.device(device)
        self.anchors = grid_anchors(spec, device=self.device)
        self.scale_factor = tuple(float(v) for v in scale_factor)

    def image(self, gid: int) -> Dict[str, object]:
        sp, dev = self.spec, self.device
        g = torch.Generator(device=dev)
        g.manual_seed(self.seed0 + int(gid))
        H, W = sp.img_hw
        lo, hi = sp.gt_range
        n_gt = int(torch.randint(lo, hi + 1, (1,), generator=g, device=dev).item())
        side = (torch.rand(n_gt, 2, generator=g, device=dev) * 0.55 + 0.05) * min(H, W)
        ctr = torch.rand(n_gt, 2, generator=g, device=dev) * torch.tensor([W, H], device=dev, dtype=torch.float32)
        x1 = (ctr[:, 0] - side[:, 0] / 2).clamp(0, W - 2)
        y1 = (ctr[:, 1] - side[:, 1] / 2).clamp(0, H - 2)
        x2 = torch.minimum(x1 + side[:, 0], torch.tensor(float(W), device=dev))
        y2 = torch.minimum(y1 + side[:, 1], torch.tensor(float(H), device=dev))
        gt = torch.stack([x1, y1, x2, y2], dim=-1)
        gt_cls = torch.randint(0, sp.num_classes, (n_gt,), generator=g, device=dev)
        c_out = sp.c_out
        cls, reg, lam = [], [], []
        for anchors, (fh, fw), a in zip(self.anchors, sp.featmaps, sp.num_anchors):
            n = anchors.shape[0]
            logit = torch.randn(n, c_out, generator=g, device=dev)
            if sp.head == HEAD_SSD:
                logit[:, -1] += 4.0
            delta = torch.randn(n, 4, generator=g, device=dev) * 0.1
            iou = _pairwise_iou(anchors, gt)
            best, arg = iou.max(dim=1)
            pos = (best >= 0.4).nonzero(as_tuple=False).squeeze(1)
            if pos.numel():
                boost = torch.rand(pos.numel(), generator=g, device=dev) * 6.0 + 2.0
                logit[pos, gt_cls[arg[pos]]] += boost
                noise = torch.randn(pos.numel(), 4, generator=g, device=dev) * 0.05
                delta[pos] = _encode(anchors[pos], gt[arg[pos]], sp.target_stds) + noise
            lmb = torch.relu(torch.randn(n, 1, generator=g, device=dev) * 0.3 + 0.5) + 1e-3
            cls.append(_to_nchw(logit, fh, fw, a))
            reg.append(_to_nchw(delta, fh, fw, a))
            lam.append(_to_nchw(lmb, fh, fw, a))
        return dict(cls=cls, reg=reg, lam=lam, gt=gt, gt_cls=gt_cls)

    def batch(self, gids: Sequence[int]) -> Dict[str, object]:
        """Head outputs for a batch: lists over levels of [B, ch, H, W] tensors + meta."""
        imgs = [self.image(g) for g in gids]
        S = self.spec.num_levels
        out = dict(
            cls_scores=[torch.stack([im["cls"][s] for im in imgs]) for s in range(S)],
            bbox_preds=[torch.stack([im["reg"][s] for im in imgs]) for s in range(S)],
            L_scores=[torch.stack([im["lam"][s] for im in imgs]) for s in range(S)],
            anchors=self.anchors,
            img_shapes=[(self.spec.img_hw[0], self.spec.img_hw[1], 3)] * len(imgs),
            scale_factors=[self.scale_factor] * len(imgs),
            gids=list(gids),
        )
        return out
