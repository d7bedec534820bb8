"""Evaluation metrics (run.py:682-711) on the device against scikit-learn on the host, same arrays.

    python tools/bench_metrics.py [--n 2000000] [--domains 30] > profiles/r2_metrics.json

The host leg is what `Run.test` + `evaluate_multi_domain` do after the epoch: `.cpu().numpy()`, one `roc_auc_score` / `log_loss`
over everything and one per domain through a pandas groupby.  The device leg is `cdcmdr_b200.metrics.evaluate_multi_domain` on
the tensors where the model left them, including the read-back of the per-domain results."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cdcmdr_b200 as cm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2_000_000)
    ap.add_argument("--domains", type=int, default=30)
    a = ap.parse_args()
    from sklearn.metrics import log_loss, roc_auc_score
    rng = np.random.default_rng(5)
    dom = rng.integers(0, a.domains, a.n).astype(np.int64)
    y = (rng.random(a.n) < 0.25).astype(np.float32)
    p = np.clip(0.25 + 0.2 * (y - 0.25) + 0.15 * rng.standard_normal(a.n), 1e-4, 1 - 1e-4).astype(np.float32)
    w = np.bincount(dom, minlength=a.domains) / a.n
    dev = torch.device("cuda", 0)
    tp, ty, td = torch.from_numpy(p).to(dev), torch.from_numpy(y).to(dev), torch.from_numpy(dom).to(dev)

    def device_leg():
        total = cm.metrics.auc_logloss(tp, ty, None, 1)
        res = cm.metrics.evaluate_multi_domain(ty, tp, td, a.domains, w)
        return total.cpu().numpy(), res
    device_leg()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        total, res = device_leg()
    torch.cuda.synchronize()
    t_dev = (time.perf_counter() - t0) / 5

    t0 = time.perf_counter()
    hp, hy, hd = tp.cpu().numpy(), ty.cpu().numpy(), td.cpu().numpy()
    auc_all, ll_all = roc_auc_score(hy, hp), log_loss(hy, hp)
    aucs = {}
    order = np.argsort(hd, kind="stable")
    cuts = np.searchsorted(hd[order], np.arange(a.domains + 1))
    for d in range(a.domains):
        idx = order[cuts[d]:cuts[d + 1]]
        aucs[d] = (roc_auc_score(hy[idx], hp[idx]), log_loss(hy[idx], hp[idx]))
    t_host = time.perf_counter() - t0
    dauc = res["domain_auc"] if isinstance(res, dict) else res[0]
    err = max(abs(dauc[d] - aucs[d][0]) for d in range(a.domains))
    print(json.dumps({"what": "per-domain + total AUC / log loss after an evaluation epoch", "samples": a.n, "domains": a.domains,
                      "device_ms": t_dev * 1e3, "host_sklearn_ms": t_host * 1e3, "speedup": t_host / t_dev,
                      "max_abs_auc_diff": float(err), "total_auc": [float(total[0, 0]), float(auc_all)],
                      "total_logloss": [float(total[0, 1]), float(ll_all)]}))


if __name__ == "__main__":
    main()
