"""The routes a kept row can take through K1 - parked by the streaming pass and read by the warp-per-row kernel
(default), parked and staged by bulk-async copies, parked and read by the thread-per-row kernel, or not parked at
all (strided gather / coalesced rescan) - must give bit-identical outputs: every route runs the same arithmetic.
The route is chosen once per process (environment), so each variant scores the same batch in a subprocess."""
import hashlib
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SNIPPET = r"""
import hashlib, json, sys
import numpy as np, torch
sys.path.insert(0, {root!r})
from aod_meh_hua_b200.scoring import Scorer
from aod_meh_hua_b200.specs import ScoringParams, get_spec
from aod_meh_hua_b200.synth import SyntheticPool
out = {{}}
for name, ids in (("cfg1_retina_r50_512_voc", [0, 1, 2]), ("cfg2_ssd300_voc", [3, 4]), ("cfg4_ssd512_coco", [5])):
    spec = get_spec(name)
    bt = SyntheticPool(spec, seed0=20, device="cpu").batch(ids)
    sc = Scorer(spec, ScoringParams(n_samples=64), max_batch=len(ids), device="cuda:0")
    r = sc.score(bt["cls_scores"], bt["bbox_preds"], bt["L_scores"], bt["anchors"], bt["img_shapes"], bt["scale_factors"],
                 image_ids=bt["gids"])
    torch.cuda.synchronize()
    h = hashlib.sha256()
    for t in (r.topk_idx, r.score_rows, r.lam_rows, r.boxes, r.row_max, r.row_argmax, r.n_det, r.pair_off,
              r.image_scores):
        h.update(t.cpu().numpy().tobytes())
    nd = r.n_det.cpu().numpy()
    h.update(b"".join(r.dets[i, :nd[i]].cpu().numpy().tobytes() + r.det_flat[i, :nd[i]].cpu().numpy().tobytes()
                      for i in range(len(ids))))
    out[name] = dict(digest=h.hexdigest(), captured=[int(v) for v in (np.asarray(sc.capture_counts()) >= 0).sum(axis=0)],
                     scores=[float(v) for v in r.image_scores.cpu()])
print("RESULT " + json.dumps(out))
"""


def _run(env_extra):
    env = dict(os.environ)
    for k in ("MEHHUA_NO_CAPTURE", "MEHHUA_NO_PARKED_KERNEL", "MEHHUA_PARKED_BULK"):
        env.pop(k, None)
    env.update(env_extra)
    p = subprocess.run([sys.executable, "-c", SNIPPET.format(root=ROOT)], env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


def test_every_row_route_gives_identical_outputs():
    base = _run({})
    assert sum(base["cfg1_retina_r50_512_voc"]["captured"]) > 0          # the default does capture on these shapes
    assert all(v > 0 for v in base["cfg1_retina_r50_512_voc"]["scores"])
    for env in ({"MEHHUA_PARKED_BULK": "1"}, {"MEHHUA_NO_PARKED_KERNEL": "1"}, {"MEHHUA_NO_CAPTURE": "1"}):
        got = _run(env)
        for name in base:
            assert got[name]["digest"] == base[name]["digest"], (env, name, got[name]["scores"], base[name]["scores"])
    assert sum(_run({"MEHHUA_NO_CAPTURE": "1"})["cfg1_retina_r50_512_voc"]["captured"]) == 0
