#!/usr/bin/env python
"""Residual flip count of the integer decisions (SURVEY 7 "hard part 1", VERDICT r1 missing #1).

    python tools/flip_study.py [--workload cfg3_retina_r50_800x1344_coco] [--images 10000] [--batch 8]
                               [--cpu-images 200] [--out gpurun_out/flip_study.txt]

Every image goes through the CUDA path (K1, K3a, K3b, K2 analytic, K3c) and through the oracle port of
the reference's own torch code running in torch-eager on the same GPU (ATen's CUDA softmax / topk /
sort - the arithmetic the reference itself has on a GPU); the first --cpu-images also through the
oracle on torch CPU (ATen's vectorised CPU softmax).  The Monte-Carlo step is replaced on all sides
by its T -> infinity limit so that image scores are comparable.  Counted per decision type:
  topk_set     priors in one side's per-level top-k but not the other's        (Lambda_L2.py:290)
  topk_order   rows whose position inside the sorted top-k differs (near-tied neighbours swapped)
  level_fg     (image, level) foreground flags                                (Lambda_L2.py:496-502)
  nms_keep     detections (flat row*C+class) kept by one side only             (bbox_nms.py:41-93)
  n_obj        images whose object count differs                               (Lambda_L2.py:344)
  obj_order    images whose objects (detections above 0.3) come in a different order (near-tied det scores)
  pairs        (row prior, object) pairs of one side only, an object being identified by the detection
               it is, not by its rank                                          (Lambda_L2.py:505-509)
  pair_cls     common pairs whose class key differs                            (Lambda_L2.py:526)
  score        images whose score differs by more than 1e-4 relative
  selected     ids in one side's top-2.5 % set but not the other's              (active_datasets.py:124)
No GPU-vs-oracle difference is hidden: a flip in an early stage is followed into the later ones.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

KEYS = ["topk_set", "topk_order", "level_fg", "nms_keep", "n_obj", "obj_order", "pairs", "pair_cls", "score"]


def compare(spec, res, out, B, counts, totals):
    S = spec.num_levels
    koff = np.concatenate([[0], np.cumsum(spec.level_k)])
    idx = res.topk_idx.cpu().numpy()
    for s in range(S):
        want = out["lvl_idx"][s].cpu().numpy()
        got = idx[:, koff[s]:koff[s + 1]]
        for b in range(B):
            totals["topk_set"] += want.shape[1]
            totals["topk_order"] += want.shape[1]
            if not np.array_equal(got[b], want[b]):
                counts["topk_set"] += len(np.setxor1d(got[b], want[b]))
                counts["topk_order"] += int((got[b] != want[b]).sum())
    fg_want = np.asarray(out["level_fg"]).astype(bool)
    counts["level_fg"] += int((res.level_fg.cpu().numpy().astype(bool) != fg_want).sum())
    totals["level_fg"] += fg_want.size
    n_det = res.n_det.cpu().numpy()
    n_obj = res.n_obj.cpu().numpy()
    poff = res.pair_off.cpu().numpy()
    # prior index of every row (rows are level-major in top-k order): pairs are compared by prior, so a
    # pure re-ordering of near-tied rows is not counted twice
    prior_got = idx
    lvl_of_row = np.repeat(np.arange(S), spec.level_k)
    prior_want = np.concatenate([out["lvl_idx"][s].cpu().numpy() for s in range(S)], axis=1)
    for b in range(B):
        dw = out["det_flat"][b].cpu().numpy()
        # det_flat is row*C + class with the side's own row numbering: translate rows to (level, prior)
        def key(flat, prior):
            nfg = spec.num_classes
            r, c = flat // nfg, flat % nfg
            return set(zip(lvl_of_row[r].tolist(), prior[r].tolist(), c.tolist()))
        kg = key(res.det_flat[b, :n_det[b]].cpu().numpy().astype(np.int64), prior_got[b])
        kw = key(dw.astype(np.int64), prior_want[b])
        counts["nms_keep"] += len(kg ^ kw)
        totals["nms_keep"] += len(kw)
        d = out["dets"][b]
        nobj_w = int((d[:, 4] > 0.3).sum())
        counts["n_obj"] += int(n_obj[b] != nobj_w)
        totals["n_obj"] += 1
        n = poff[b, S]
        rg = res.pair_row[b, :n].cpu().numpy()
        og = res.pair_obj[b, :n].cpu().numpy()
        cg = res.pair_cls[b, :n].cpu().numpy()
        # an object is identified by the detection it is (level, prior, class), not by its rank among the
        # detections: two near-tied detection scores may swap ranks without changing any membership
        def det_ids(flat, prior):
            nfg = spec.num_classes
            return [(int(lvl_of_row[f // nfg]), int(prior[f // nfg]), int(f % nfg)) for f in flat]
        ids_g = det_ids(res.det_flat[b, :n_det[b]].cpu().numpy().astype(np.int64), prior_got[b])
        ids_w = det_ids(dw.astype(np.int64), prior_want[b])
        counts["obj_order"] += int(ids_g[:nobj_w] != ids_w[:nobj_w])
        totals["obj_order"] += 1
        pg = {(int(lvl_of_row[r]), int(prior_got[b][r]), ids_g[o]): int(c) for r, o, c in zip(rg, og, cg)}
        pw = {}
        for rec in out["flat"]:
            if rec["image"] != b:
                continue
            for r, o, c in zip(rec["row"], rec["obj"], rec["cls"]):
                pw[(int(lvl_of_row[r]), int(prior_want[b][r]), ids_w[o])] = int(c)
        counts["pairs"] += len(set(pg) ^ set(pw))
        totals["pairs"] += len(pw)
        common = set(pg) & set(pw)
        counts["pair_cls"] += sum(1 for k in common if pg[k] != pw[k])
        totals["pair_cls"] += len(common)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3_retina_r50_800x1344_coco")
    ap.add_argument("--images", type=int, default=10000)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--cpu-images", type=int, default=200)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "flip_study.txt"))
    args = ap.parse_args()
    from aod_meh_hua_b200.scoring import Scorer
    from aod_meh_hua_b200.specs import ScoringParams, get_spec
    from aod_meh_hua_b200.synth import SyntheticPool
    from oracle import meh_hua_oracle as O
    spec = get_spec(args.workload)
    params = ScoringParams(n_samples=0)
    dev = torch.device("cuda:0")
    synth = SyntheticPool(spec, seed0=20, device=dev)
    sc = Scorer(spec, params, max_batch=args.batch, device=dev)
    kw = O.spec_kwargs(spec, params)
    kw.pop("T")
    arms = {"oracle on torch CUDA": dict(n=args.images, cpu=False), "oracle on torch CPU": dict(n=args.cpu_images, cpu=True)}
    lines = [f"# flip study: {args.workload}, {args.images} images (ids 0..{args.images - 1}, seed0 20), batch {args.batch}; "
             f"analytic (T -> infinity) uncertainty on every side",
             f"# torch {torch.__version__}, {torch.cuda.get_device_name(0)}"]
    for arm, cfg in arms.items():
        if cfg["n"] <= 0:
            continue
        counts = {k: 0 for k in KEYS}
        totals = {k: 0 for k in KEYS}
        got_all, want_all = [], []
        t0 = time.time()
        for i0 in range(0, cfg["n"], args.batch):
            gids = list(range(i0, min(i0 + args.batch, cfg["n"])))
            batch = synth.batch(gids)
            res = sc.score(batch["cls_scores"], batch["bbox_preds"], batch["L_scores"], batch["anchors"],
                           batch["img_shapes"], batch["scale_factors"], image_ids=gids)
            if cfg["cpu"]:
                ob = {k: ([t.cpu() for t in v] if isinstance(v, list) and torch.is_tensor(v[0]) else v) for k, v in batch.items()}
            else:
                ob = batch
            out = O.score_batch(ob, analytic=True, record_flat=True, **kw)
            compare(spec, res, out, len(gids), counts, totals)
            got_all.append(res.image_scores.cpu().numpy().copy())
            want_all.append(np.asarray([float(v) for v in out["image_scores"]], dtype=np.float32))
        got, want = np.concatenate(got_all), np.concatenate(want_all)
        rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-12)
        rel[(got == 0) & (want == 0)] = 0
        counts["score"] = int((rel > 1e-4).sum())
        totals["score"] = len(got)
        k = max(1, int(0.025 * len(got)))
        sel_g = set(np.argsort(got, kind="stable")[-k:].tolist())
        sel_w = set(np.argsort(want, kind="stable")[-k:].tolist())
        lines.append(f"\n## CUDA path vs {arm}: {cfg['n']} images, {time.time() - t0:.0f} s")
        lines.append(f"{'decision':12s} {'differing':>10s} {'of':>12s}")
        for key in KEYS:
            lines.append(f"{key:12s} {counts[key]:10d} {totals[key]:12d}")
        lines.append(f"{'selected':12s} {len(sel_g ^ sel_w):10d} {k:12d}   (top 2.5 % of the {len(got)} images)")
        lines.append(f"max relative score difference among the others: {float(rel[rel <= 1e-4].max()):.2e}; "
                     f"images with score 0 on one side only: {int(((got == 0) != (want == 0)).sum())}")
        worst = np.argsort(rel)[::-1][:5]
        lines.append("largest score differences (image id, cuda path, oracle): " +
                     ", ".join(f"({int(i)}, {got[i]:.6f}, {want[i]:.6f})" for i in worst if rel[i] > 1e-4))
        print("\n".join(lines[-14:]), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
