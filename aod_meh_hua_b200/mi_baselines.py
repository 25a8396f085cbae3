"""Mutual-information scoring of the reference's two competing baselines, on the GPU (kernel KM).

Reference: mmdet/apis/CalEnsembleUnc.py:137-181 (`Ensemble_MI`, `ComputeMI`: three independently trained
detectors) and mmdet/apis/CalMCDropoutUnc.py:137-201 (`MCDropout_MI`, `ComputeMCDropoutMI`: n = 25
stochastic passes of one detector).  Both turn the members' raw classification maps of a batch - per
level [B, A*nCls, H, W] - into one score per image: sigmoid, mean over members, entropy of the mean minus
the mean entropy, averaged over the level's priors and then over the levels.  Here that is one streaming
pass over the members' logits (`mehhua_mi_score_batch`); the function names, arguments and return types
are the reference's.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import torch

from . import _lib


def mutual_information(members: Sequence[Sequence[torch.Tensor]], n_cls: int):
    """members[m][s] = member m's level-s classification map [B, A*n_cls, H, W] (CUDA, fp32).
    Returns (image_scores float32[B], level_mi float32[B, S]) on the same device."""
    lib = _lib.load()
    M, S = len(members), len(members[0])
    if M < 1 or M > 32 or S < 1 or S > _lib.MAX_LEVELS:
        raise _lib.MehhuaError(f"mutual_information: {M} members x {S} levels (at most 32 x {_lib.MAX_LEVELS})")
    first = members[0][0]
    if not first.is_cuda:
        raise _lib.MehhuaError("mutual_information needs CUDA tensors; there is no CPU fallback")
    dev, B = first.device, int(first.shape[0])
    lv = _lib.LevelArray()
    keep: List[torch.Tensor] = []
    ptrs = (C.c_void_p * (M * S))()
    for s in range(S):
        _, ch, H, W = members[0][s].shape
        if ch % n_cls:
            raise _lib.MehhuaError(f"level {s}: {ch} channels are not a multiple of nCls = {n_cls}")
        lv[s].H, lv[s].W, lv[s].A = int(H), int(W), int(ch // n_cls)
        for m in range(M):
            t = members[m][s]
            if tuple(t.shape) != (B, ch, H, W) or t.device != dev:
                raise _lib.MehhuaError(f"member {m}, level {s}: shape / device mismatch")
            t = t.detach().float().contiguous()
            keep.append(t)
            ptrs[m * S + s] = t.data_ptr()
    ws_bytes = int(lib.mehhua_mi_workspace_bytes(lv, S, B))
    ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
    level_mi = torch.empty(B, S, device=dev)
    scores = torch.empty(B, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _lib.check(lib.mehhua_mi_score_batch(ptrs, M, lv, S, int(n_cls), B, level_mi.data_ptr(), scores.data_ptr(),
                                             ws.data_ptr(), ws.numel(), st), "mehhua_mi_score_batch")
    return scores, level_mi


def _as_levels(out):
    """One member's `justOut=True` forward: a list over levels of [B, A*nCls, H, W] maps (a tensor, or a list of
    per-image [A*nCls, H, W] maps as the reference iterates them)."""
    return [o if torch.is_tensor(o) else torch.stack(list(o)) for o in out]


def ComputeMI(m1Out, m2Out, m3Out, nCls=20):
    """CalEnsembleUnc.py:166-181: list of B python floats."""
    scores, _ = mutual_information([_as_levels(m1Out), _as_levels(m2Out), _as_levels(m3Out)], nCls)
    return scores.tolist()


def ComputeMCDropoutMI(*MCOuts, nCls=20):
    """CalMCDropoutUnc.py:185-201: list of B python floats from n stochastic passes."""
    scores, _ = mutual_information([_as_levels(o) for o in MCOuts], nCls)
    return scores.tolist()


def _unwrap(data):
    data = dict(data)
    data["img"] = getattr(data["img"], "data", data["img"])
    data["img_metas"] = getattr(data["img_metas"], "data", data["img_metas"])
    return data


def Ensemble_MI(m1, m2, m3, data_loader, nCls=20, **kwargs):
    """CalEnsembleUnc.py:137-164: the pool loop of the three-member ensemble; returns a float tensor [pool]."""
    for m in (m1, m2, m3):
        m.eval()
    uncertainties = []
    with torch.no_grad():
        for i, data in enumerate(data_loader):
            data = _unwrap(data)
            outs = [m(return_loss=False, rescale=True, isEval=True, justOut=True, batchIdx=i, **data, **kwargs)
                    for m in (m1, m2, m3)]
            uncertainties.extend(ComputeMI(*outs, nCls=nCls))
    return torch.tensor(uncertainties)


def activate_dropout(model):
    """The part of utils/functions.py:500-505 that is not model surgery: every nn.Dropout2d goes back to train
    mode (the reference also inserts Dropout2d modules after ReLUs with `append_dropout`, :491-498 - that
    belongs to the model, not to the scoring path)."""
    for module in model.modules():
        if isinstance(module, torch.nn.Dropout2d):
            module.train()


def MCDropout_MI(model, data_loader, n=25, nCls=20, activate=activate_dropout, **kwargs):
    """CalMCDropoutUnc.py:137-165: n stochastic passes per batch; returns a float tensor [pool]."""
    model.eval()
    activate(model)
    uncertainties = []
    with torch.no_grad():
        for i, data in enumerate(data_loader):
            data = _unwrap(data)
            outs = [model(return_loss=False, rescale=True, isEval=True, justOut=True, batchIdx=i, **data, **kwargs)
                    for _ in range(n)]
            uncertainties.extend(ComputeMCDropoutMI(*outs, nCls=nCls))
    return torch.tensor(uncertainties)
