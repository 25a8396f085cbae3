"""Pool-level plumbing: sharding the unlabelled pool by image over ranks, the single all-gather of
per-image scores, and selection (update_X_L with its deterministic top-k part on the GPU, K4).

Reference: tools/train_RetinaNet.py:221-251 (scoring stage of the AL cycle, single process,
un-sharded), mmdet/utils/active_datasets.py:102-135 (update_X_L).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of image ids owned by `rank`: the first n % world ranks get one extra."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_scores(local: torch.Tensor, n_total: int, rank: int, world: int, group=None) -> torch.Tensor:
    """All ranks' score shards -> float32[n_total] on every rank (one all_gather; shards are padded
    to the largest shard so the collective is regular).  Works on the tensor's own device: NCCL
    for CUDA tensors, gloo for CPU tensors (the CPU tests)."""
    import torch.distributed as dist
    start, stop = shard_range(n_total, rank, world)
    if local.numel() != stop - start:
        raise ValueError(f"rank {rank} holds {local.numel()} scores, expected {stop - start}")
    if world == 1:
        return local.float().clone()
    width = (n_total + world - 1) // world
    pad = torch.zeros(width, dtype=torch.float32, device=local.device)
    pad[: local.numel()] = local.float()
    out = torch.empty(world * width, dtype=torch.float32, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    parts = []
    for r in range(world):
        a, b = shard_range(n_total, r, world)
        parts.append(out[r * width: r * width + (b - a)])
    return torch.cat(parts)


def select_top(uncertainty: torch.Tensor, candidate_mask: Optional[torch.Tensor], n_top: int) -> torch.Tensor:
    """ids of the n_top largest scores among candidates, on the GPU (K4)."""
    from .scoring import pool_topk
    return pool_topk(uncertainty, n_top, candidate_mask)


def update_X_L(uncertainty, X_all, X_L, X_S_size, device=None, **kwargs):
    """Drop-in for mmdet.utils.active_datasets.update_X_L (:102-135): same arguments, same return
    (sorted X_L_next, X_U_next).  The arg[-n:] part - the global top-k - runs on the GPU; the two
    host-RNG draws (zero-score picks, X_U shuffle) use numpy's global state exactly as the
    reference does.  Ties in the score go to the larger image id (the reference's unstable argsort
    leaves them unpinned).  device: where K4 runs - default: the device of `uncertainty` when it is a
    CUDA tensor, else the process's current CUDA device (rank r's GPU under torchrun)."""
    if device is None:
        if torch.is_tensor(uncertainty) and uncertainty.is_cuda:
            device = uncertainty.device
        else:
            device = torch.device("cuda", torch.cuda.current_device())
    if torch.is_tensor(uncertainty):
        unc_t = uncertainty.detach().float()
        uncertainty = unc_t.cpu().numpy()
    else:
        uncertainty = np.asarray(uncertainty)
        unc_t = torch.from_numpy(uncertainty.astype(np.float32))
    X_all = np.asarray(X_all)
    X_L = np.asarray(X_L)
    all_X_U = np.array(list(set(X_all.tolist()) - set(X_L.tolist())))
    mask = torch.zeros(unc_t.numel(), dtype=torch.uint8)
    mask[torch.from_numpy(all_X_U.astype(np.int64))] = 1
    unc_dev = unc_t.to(device)
    mask_dev = mask.to(device)
    if kwargs.get("zeroRate"):
        u = uncertainty[all_X_U]
        zeros = (u == 0).nonzero()[0]
        zero_size = int(X_S_size * kwargs["zeroRate"])
        non_zero_size = X_S_size - zero_size
        zero_size = min(zero_size, len(zeros))
        mode = kwargs.get("useMaxConf", "False")
        if mode != "False":
            order = np.array(kwargs["maxconf"])[all_X_U].argsort()
            zero_idx = order[:zero_size] if mode == "min" else order[-zero_size:]
        else:
            zero_idx = np.random.choice(zeros, zero_size)
        top = select_top(unc_dev, mask_dev, non_zero_size).cpu().numpy()
        X_S = np.concatenate((all_X_U[zero_idx], top[::-1]))
    else:
        X_S = select_top(unc_dev, mask_dev, X_S_size).cpu().numpy()[::-1]
    X_L_next = np.concatenate((X_L, X_S))
    rest = np.array(list(set(X_all.tolist()) - set(X_L_next.tolist())))
    np.random.shuffle(rest)
    X_U_next = rest[: X_L_next.shape[0]]
    X_L_next.sort()
    X_U_next.sort()
    return X_L_next, X_U_next


# ---------------------------------------------------------------------------------------------
# On-disk formats of the active-learning cycle (the caller side of the path)
# ---------------------------------------------------------------------------------------------
def save_cycle(work_dir: str, cycle: int, X_L, X_U, uncertainty) -> None:
    """What tools/train_RetinaNet.py:249-251 writes after scoring cycle `cycle`: the next cycle's
    labelled / unlabelled index sets and the score vector, as plain .npy files named
    X_L_{cycle+1}.npy, X_U_{cycle+1}.npy, Unc_{cycle+1}.npy (dtype and shape as returned by
    update_X_L / calculate_uncertainty: int64 index arrays, float32[N_all])."""
    if torch.is_tensor(uncertainty):
        uncertainty = uncertainty.detach().cpu().numpy()
    elif not isinstance(uncertainty, np.ndarray):
        uncertainty = torch.stack(list(uncertainty)).numpy()        # list of 0-d tensors (:242-245)
    np.save(f"{work_dir}/X_L_{cycle + 1}.npy", np.asarray(X_L))
    np.save(f"{work_dir}/X_U_{cycle + 1}.npy", np.asarray(X_U))
    np.save(f"{work_dir}/Unc_{cycle + 1}.npy", uncertainty)


def ResumeCycle_WorkDir(work_dir: str, currentCycle: int, fromStartCycle: int):
    """mmdet/utils/functions.py:485-490: (X_L, X_U) saved for `fromStartCycle`, or (False, False)
    while the loop is still before that cycle."""
    if currentCycle < fromStartCycle:
        return (False, False)
    return (np.load(f"{work_dir}/X_L_{fromStartCycle}.npy"), np.load(f"{work_dir}/X_U_{fromStartCycle}.npy"))


def ResumeCycle(cfg, currentCycle: int, fromStartCycle: int):
    """mmdet/utils/functions.py:478-483 (same, work_dir taken from the config)."""
    return ResumeCycle_WorkDir(cfg.work_dir, currentCycle, fromStartCycle)


def scoring_stage(cfg, poolModel, data_loader, X_all, X_L, cycle: int, *, zeroRate=0.15, useMaxConf="False",
                  saveMaxConf=False, clsW=False, scaleUnc=False, score_thr=0.3, iou_thr=0.9, save=True,
                  calculate=None):
    """The scoring stage of one AL cycle exactly as the training scripts run it
    (tools/train_RetinaNet.py:231-251): calculate_uncertainty over the pool loader -> update_X_L ->
    X_L / X_U / Unc .npy files.  Returns (X_L_next, X_U_next, uncertainty ndarray)."""
    if calculate is None:
        from .dropin import calculate_uncertainty as calculate
    with torch.no_grad():
        out = calculate(cfg, poolModel, data_loader, return_box=False, showNMS=False, saveUnc=False,
                        saveMaxConf=saveMaxConf, clsW=clsW, scaleUnc=scaleUnc, score_thr=score_thr, iou_thr=iou_thr)
    maxconf, uncertainty = (out[1], out[0]) if saveMaxConf else (None, out)
    if torch.is_tensor(uncertainty):
        uncertainty = uncertainty.numpy()
    elif not isinstance(uncertainty, np.ndarray):
        uncertainty = torch.stack(uncertainty).numpy()
    X_L_next, X_U_next = update_X_L(uncertainty, X_all, X_L, cfg.X_S_size, zeroRate=zeroRate, maxconf=maxconf,
                                    useMaxConf=useMaxConf)
    if save:
        save_cycle(cfg.work_dir, cycle, X_L_next, X_U_next, uncertainty)
    return X_L_next, X_U_next, uncertainty
