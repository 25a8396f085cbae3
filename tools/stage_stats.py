#!/usr/bin/env python
"""Print data-dependent sizes of the scoring path for a workload (candidates, objects, pairs per
image) - used to size the kernels and to state the per-launch work in DESIGN.md / profiles."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from aod_meh_hua_b200.scoring import Scorer  # noqa: E402
from aod_meh_hua_b200.specs import ScoringParams, get_spec  # noqa: E402
from aod_meh_hua_b200.synth import SyntheticPool  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg3_retina_r50_800x1344_coco")
ap.add_argument("--images", type=int, default=16)
ap.add_argument("--samples", type=int, default=500)
a = ap.parse_args()
spec = get_spec(a.workload)
pool = SyntheticPool(spec, seed0=20, device="cuda:0")
sc = Scorer(spec, ScoringParams(n_samples=a.samples), max_batch=a.images, device="cuda:0")
batch = pool.batch(list(range(a.images)))
res = sc.score(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"], batch["img_shapes"],
               batch["scale_factors"], image_ids=batch["gids"])
torch.cuda.synchronize()
S = spec.num_levels
pairs = res.pair_off[:, S].cpu()
print("workload", spec.name, "images", a.images, "status", sc.read_status())
print("n_det  ", res.n_det.cpu().tolist())
print("n_obj  ", res.n_obj.cpu().tolist())
print("pairs  ", pairs.tolist(), "mean", float(pairs.float().mean()))
print("pairs/level mean", (res.pair_off[:, 1:] - res.pair_off[:, :-1]).float().mean(0).cpu().tolist())
print("level_fg", res.level_fg.cpu().float().mean(0).tolist())
print("fg rows (row_max > 0.3) per image", (res.row_max > 0.3).sum(1).cpu().tolist())
print("scores ", [round(float(v), 4) for v in res.image_scores.cpu()])
print("draws per image (T*C*P) mean", float(pairs.float().mean()) * a.samples * spec.c_out)
