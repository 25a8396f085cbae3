#!/usr/bin/env python
"""Rows parked per (image, level) by K1's capture mode on a batch of synthetic images, and how many
(image, level) estimates missed (fell back to the select over all keys + gather):
    python tools/capture_stats.py [--workload NAME] [--images N]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from aod_meh_hua_b200.scoring import Scorer  # noqa: E402
from aod_meh_hua_b200.specs import ScoringParams, get_spec  # noqa: E402
from aod_meh_hua_b200.synth import SyntheticPool  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg3_retina_r50_800x1344_coco")
ap.add_argument("--images", type=int, default=296)
ap.add_argument("--chunk", type=int, default=74)
a = ap.parse_args()
spec = get_spec(a.workload)
pool = SyntheticPool(spec, seed0=20, device="cuda:0")
sc = Scorer(spec, ScoringParams(), max_batch=a.chunk, device="cuda:0")
rows = []
for g0 in range(0, a.images, a.chunk):
    ids = list(range(g0, min(g0 + a.chunk, a.images)))
    bt = pool.batch(ids)
    sc.bind(bt["cls_scores"], bt["bbox_preds"], bt["L_scores"], bt["anchors"], bt["img_shapes"], bt["scale_factors"],
            image_ids=bt["gids"])
    sc.k1()
    torch.cuda.synchronize()
    rows.append(np.asarray(sc.capture_counts()).reshape(len(ids), spec.num_levels))
cnt = np.concatenate(rows)
out = dict(workload=spec.name, images=int(cnt.shape[0]), levels=[])
for s in range(spec.num_levels):
    c = cnt[:, s]
    if (c < 0).all():
        out["levels"].append(dict(level=s, capture=False))
        continue
    k = min(spec.nms_pre, spec.level_priors(s)) if hasattr(spec, "level_priors") else spec.nms_pre
    out["levels"].append(dict(level=s, capture=True, parked_min=int(c.min()), parked_mean=float(c.mean()), parked_max=int(c.max()),
                              missed=int(((c < spec.nms_pre) | (c > 4096)).sum())))
print(json.dumps(out))
