#!/usr/bin/env python
"""Golden fixtures for CDC training steps whose domain (hence the selected tower) CHANGES from step to step, produced by the
UNMODIFIED reference:   python tests/golden/make_golden_alternate.py

`model(x, mode='split', domain_i=d)` selects one tower column; the parameters that neither that column nor the regulariser
reaches keep `.grad is None` and torch.optim.Adam SKIPS them (no moment decay, no step).  A tower that was trained at step s
and is idle at step s+1 therefore must not coast on its Adam momentum - something two identical steps cannot show.  Five steps
with domains [1, 0, 2, 1, 3] -> towers [2, 0, 1, 2, 2] on CDC(PLE / MMoE / STAR); same layout as make_golden.py."""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (installs the import shim, imports the reference models)

DOMAINS = [1, 0, 2, 1, 3]


def main():
    torch.manual_seed(2001)
    torch.set_num_threads(1)
    rng = np.random.default_rng(2001)
    B, T = 24, 3
    batches = [G.make_batch(rng, B, T) for _ in range(len(DOMAINS))]
    cfg = G.Cfg()
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    try:
        n_domain = int(G.FIELD_DIMS[3])
        d2g = [0, 2, 1, 2]
        w = [0.4, 0.3, 0.2, 0.1]
        for base, ed, td in (("ple", ((16, 8), (8,)), (8, 4)), ("mmoe", (16, 8), (8, 4)), ("star", None, (16, 8))):
            c = G.CDC(G.FIELD_DIMS, G.E, T, n_domain, base, ed, td, 3, domain_cnt_weight=w, n_causal_mask=5,
                      device="cpu", dropout=0.0, config=cfg, **G.L2)
            c.domain2group_list = list(d2g)
            c.domain2group = torch.tensor(d2g, dtype=torch.int64)
            out = {"d2g": np.array(d2g), "domains": np.array(DOMAINS)}
            for k, v in G.sd_np(c).items():
                out["sd0." + k] = v
            opt = torch.optim.Adam(c.parameters(), lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)
            crit = torch.nn.BCELoss()
            c.train()
            for s, dom in enumerate(DOMAINS):
                x, y, g = batches[s]
                out[f"in{s}.x"], out[f"in{s}.y"] = x, y
                p = c(torch.from_numpy(x), mode="split", domain_i=dom)
                bce = crit(p, torch.from_numpy(y).squeeze().float())
                reg = c.get_regularization_loss(device="cpu")
                c.zero_grad()
                (bce + reg).backward()
                out[f"step{s}.pred"] = p.detach().numpy().copy()
                out[f"step{s}.bce"], out[f"step{s}.reg"] = np.float32(bce.item()), np.float32(reg.item())
                out[f"step{s}.none"] = np.array(sorted(k for k, q in c.named_parameters() if q.grad is None))
                opt.step()
                for k, v in G.sd_np(c).items():
                    out[f"sd{s + 1}." + k] = v
            path = os.path.join(HERE, f"cdc_{base}_alternate.npz")
            np.savez_compressed(path, **out)
            print(base, os.path.getsize(path) // 1024, "KiB")
    finally:
        os.chdir(cwd)


if __name__ == "__main__":
    main()
