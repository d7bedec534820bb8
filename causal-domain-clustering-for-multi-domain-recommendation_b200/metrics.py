"""Evaluation metrics of the reference's runner on the device (reference run.py:647-711).

`Run.test` moves every batch's predictions / labels / domain ids to the host (`.cpu().numpy()` per batch) and calls scikit-learn's
`roc_auc_score` / `log_loss` on the concatenation - once over everything, once per domain (`evaluate_multi_domain`).  Here the
batches stay where the model left them: `DeviceEvaluator.add()` keeps device tensors, `result()` runs ONE sort by (domain,
prediction) and a fixed-order reduction per domain (cdcmdr_auc_logloss) and reads back 4 doubles per domain.

    ev = cm.metrics.DeviceEvaluator(n_domain, domain_cnt_weight)
    for X, y in batches:  ev.add(model(X, ...), y, X[:, domain_idx])
    result_dict = ev.result()            # the keys run.py:682-711 produces: total_auc, total_loss, domain_auc, domain_loss, mean_*
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def auc_logloss(pred: torch.Tensor, target: torch.Tensor, domain: torch.Tensor | None, n_domain: int) -> torch.Tensor:
    """-> float64 [n_domain, 4] = (AUC, log loss, positives, samples) per domain on pred's device (domain None: one set)."""
    lib = _lib.load()
    pred = pred.reshape(-1).to(torch.float32).contiguous()
    n = pred.numel()
    target = target.reshape(-1).contiguous()
    if target.dtype not in (torch.int16, torch.float32):
        target = target.to(torch.float32)
    if target.numel() != n or (domain is not None and domain.numel() != n):
        raise ValueError("cdcmdr.metrics: predictions, targets and domains must have the same length")
    if domain is not None:
        domain = domain.reshape(-1).contiguous()
        if domain.dtype not in (torch.int32, torch.int64):
            domain = domain.to(torch.int64)
    emu = getattr(lib, "is_host_emulator", False)
    if pred.device.type != "cuda" and not emu:
        raise RuntimeError("cdcmdr.metrics runs on a CUDA device (sm_100a); there is no CPU fallback")
    out = torch.empty(n_domain, 4, dtype=torch.float64, device=pred.device)
    scratch = torch.empty(max(int(lib.auc_logloss_scratch_bytes(n, n_domain)), 256), dtype=torch.uint8, device=pred.device)
    stream = torch.cuda.current_stream(pred.device).cuda_stream if pred.device.type == "cuda" else 0
    lib.auc_logloss(pred.data_ptr(), target.data_ptr(), 1 if target.dtype == torch.float32 else 0,
                    domain.data_ptr() if domain is not None else None, 1 if (domain is not None and domain.dtype == torch.int64) else 0,
                    n, n_domain, out.data_ptr(), scratch.data_ptr(), stream)
    return out


def evaluate_multi_domain(targets, predicts, domains, n_domain, domain_cnt_weight, return_type="dict"):
    """run.py:690-711 on device tensors: per-domain AUC / log loss (NaN where a domain holds one class only; domains without
    samples are absent from the dicts, as pandas' groupby leaves them out) and their domain_cnt_weight-weighted means."""
    res = auc_logloss(predicts, targets, domains, n_domain).cpu().numpy()
    if return_type == "dict":
        domain_auc, domain_loss = dict(), dict()
    else:
        domain_auc, domain_loss = np.zeros(n_domain), np.zeros(n_domain)
    mean_auc, mean_loss = 0, 0
    for d in range(n_domain):
        if res[d, 3] == 0:
            continue
        auc, loss = float(res[d, 0]), float(res[d, 1])
        domain_auc[d], domain_loss[d] = auc, loss
        mean_auc += domain_cnt_weight[d] * auc
        mean_loss += domain_cnt_weight[d] * loss
    return dict({'domain_auc': domain_auc, 'domain_loss': domain_loss, 'mean_auc': mean_auc, 'mean_loss': mean_loss})


class DeviceEvaluator:
    """Accumulates one evaluation pass on the device and produces Run.test's result_dict (run.py:677-688)."""

    def __init__(self, n_domain, domain_cnt_weight=None, is_evaluate_multi_domain=True):
        self.n_domain = int(n_domain)
        self.w = [1.0 / n_domain] * n_domain if domain_cnt_weight is None else [float(v) for v in domain_cnt_weight]
        self.multi = bool(is_evaluate_multi_domain)
        self.p, self.y, self.d = [], [], []

    def add(self, pred, target, domain):
        self.p.append(pred.detach().reshape(-1))
        self.y.append(target.detach().reshape(-1))
        self.d.append(domain.detach().reshape(-1))

    def result(self):
        p, y, d = torch.cat(self.p), torch.cat(self.y), torch.cat(self.d)
        tot = auc_logloss(p, y, None, 1).cpu().numpy()[0]
        out = dict(total_auc=float(tot[0]), total_loss=float(tot[1]))
        if self.multi:
            out.update(evaluate_multi_domain(y, p, d, self.n_domain, self.w))
        return out
