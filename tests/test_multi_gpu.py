"""GPU (needs >= 2 devices; skipped otherwise): the sharded pool path over NCCL - every rank scores
its contiguous shard with Philox keys by GLOBAL image id, one all_gather, K4 on every rank.  Scores
and the selected set must be bit-identical to the single-process run."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from aod_meh_hua_b200.pool import gather_scores, select_top, shard_range
from aod_meh_hua_b200.scoring import Scorer
from aod_meh_hua_b200.specs import ScoringParams, get_spec
from aod_meh_hua_b200.synth import SyntheticPool
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
if world > 1:
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
spec = get_spec("tiny_retina_coco")
n_pool, batch = 37, 4                      # ragged: shards of 19 + 18, last batches partial
pool = SyntheticPool(spec, seed0=20, device="cpu")
sc = Scorer(spec, ScoringParams(n_samples=64), max_batch=batch, device=dev)
a, b = shard_range(n_pool, rank, world)
local = []
for lo in range(a, b, batch):
    gids = list(range(lo, min(lo + batch, b)))
    bt = pool.batch(gids)
    res = sc.score(bt["cls_scores"], bt["bbox_preds"], bt["L_scores"], bt["anchors"], bt["img_shapes"],
                   bt["scale_factors"], image_ids=gids)
    local.append(res.image_scores.clone())
local = torch.cat(local) if local else torch.zeros(0, device=dev)
full = gather_scores(local, n_pool, rank, world)
mask = torch.ones(n_pool, dtype=torch.uint8, device=dev)
mask[::5] = 0                               # "already labelled"
top = select_top(full, mask, 9)
np.save({out!r} + f"/scores_w{{world}}_r{{rank}}.npy", full.cpu().numpy())
np.save({out!r} + f"/top_w{{world}}_r{{rank}}.npy", top.cpu().numpy())
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
print("ok", rank)
"""


def _launch(tmp_path, world):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / f"worker_w{world}.py"
    script.write_text(_WORKER.format(root=ROOT, port=port, out=str(tmp_path)))
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=600)
        assert p.returncode == 0, out
        assert "ok" in out


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run with gpurun --gpus 2)")
def test_world_size_independence_nccl(tmp_path):
    _launch(tmp_path, 1)
    _launch(tmp_path, 2)
    s1 = np.load(tmp_path / "scores_w1_r0.npy")
    t1 = np.load(tmp_path / "top_w1_r0.npy")
    assert s1.shape == (37,) and (s1 > 0).any()
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"scores_w2_r{r}.npy"), s1)      # bit-identical scores
        assert np.array_equal(np.load(tmp_path / f"top_w2_r{r}.npy"), t1)         # identical selection, same order
