"""Candidate-box (anchor) grids for the two heads on the scoring path.

The anchors are an *input* of the hot path (they define the candidate-box layout the logits are
indexed by): row n of level s is location (h, w), base anchor a with n = (h*W + w)*A + a.

Semantics follow mmdet/core/anchor/anchor_generator.py (reference):
  :130-193  base anchors: w = h = base_size, h_ratio = sqrt(r), w_ratio = 1/h_ratio,
            scale_major -> ratio index outer / scale inner, else scale outer / ratio inner
  :337-380  grid: shift_x = w*stride_w, shift_y = h*stride_h, row-major, anchor-minor
  :476-564  SSD: min/max sizes from basesize_ratio_range, scales (1, sqrt(max/min)),
            ratios (1, 1/r, r ...), centers = stride/2, keep [ratio_0..] of scale 0 plus the
            ratio-1 anchor of scale 1 inserted at position 1
All arithmetic is float32 torch on the CPU so the values are the ones the reference would hold.
"""
from __future__ import annotations

from typing import List

import numpy as np
import torch

from .specs import HEAD_RETINA, HEAD_SSD, DetectorSpec


def _base_anchors(base_size: float, scales: torch.Tensor, ratios: torch.Tensor,
                  center, scale_major: bool) -> torch.Tensor:
    w = h = base_size
    x_c, y_c = center
    h_r = torch.sqrt(ratios)
    w_r = 1 / h_r
    if scale_major:
        ws = (w * w_r[:, None] * scales[None, :]).view(-1)
        hs = (h * h_r[:, None] * scales[None, :]).view(-1)
    else:
        ws = (w * scales[:, None] * w_r[None, :]).view(-1)
        hs = (h * scales[:, None] * h_r[None, :]).view(-1)
    return torch.stack([x_c - 0.5 * ws, y_c - 0.5 * hs, x_c + 0.5 * ws, y_c + 0.5 * hs], dim=-1)


def retina_base_anchors(spec: DetectorSpec) -> List[torch.Tensor]:
    n = spec.retina_scales_per_octave
    octave = np.array([2 ** (i / n) for i in range(n)])
    scales = torch.Tensor(octave * spec.retina_octave_base_scale)
    ratios = torch.Tensor(list(spec.retina_ratios))
    return [_base_anchors(float(s), scales, ratios, (0.0, 0.0), True) for s in spec.strides]


def ssd_base_anchors(spec: DetectorSpec) -> List[torch.Tensor]:
    size = spec.ssd_input_size
    lo, hi = spec.ssd_ratio_range
    lo_i, hi_i = int(lo * 100), int(hi * 100)
    step = int(np.floor(hi_i - lo_i) / (spec.num_levels - 2))
    mins, maxs = [], []
    for r in range(lo_i, hi_i + 1, step):
        mins.append(int(size * r / 100))
        maxs.append(int(size * (r + step) / 100))
    first = {(300, 0.15): (7, 15), (300, 0.2): (10, 20), (512, 0.1): (4, 10), (512, 0.15): (7, 15)}
    if (size, lo) not in first:
        raise ValueError(f"unsupported SSD size/ratio-range combination {(size, lo)}")
    a, b = first[(size, lo)]
    mins.insert(0, int(size * a / 100))
    maxs.insert(0, int(size * b / 100))
    out = []
    for k, stride in enumerate(spec.strides):
        scales = torch.Tensor([1.0, np.sqrt(maxs[k] / mins[k])])
        rl = [1.0]
        for r in spec.ssd_ratios[k]:
            rl += [1 / r, r]
        ratios = torch.Tensor(rl)
        base = _base_anchors(mins[k], scales, ratios, (stride / 2.0, stride / 2.0), False)
        keep = list(range(len(rl)))
        keep.insert(1, len(keep))
        out.append(base.index_select(0, torch.LongTensor(keep)))
    return out


def base_anchors(spec: DetectorSpec) -> List[torch.Tensor]:
    if spec.head == HEAD_RETINA:
        return retina_base_anchors(spec)
    if spec.head == HEAD_SSD:
        return ssd_base_anchors(spec)
    raise ValueError("unknown head")


def grid_anchors(spec: DetectorSpec, device="cpu") -> List[torch.Tensor]:
    """[N_s, 4] float32 per level, n = (h*W + w)*A + a."""
    out = []
    for base, (fh, fw), stride in zip(base_anchors(spec), spec.featmaps, spec.strides):
        sx = torch.arange(0, fw) * stride
        sy = torch.arange(0, fh) * stride
        xx = sx.repeat(fh)
        yy = sy.view(-1, 1).repeat(1, fw).view(-1)
        shifts = torch.stack([xx, yy, xx, yy], dim=-1).type_as(base)
        out.append((base[None, :, :] + shifts[:, None, :]).view(-1, 4).contiguous().to(device))
    return out
