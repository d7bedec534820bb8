#!/usr/bin/env python
"""Golden fixtures for the CDC affinity-matrix probing loop, produced by the UNMODIFIED reference:

    python tests/golden/make_golden_probe.py

`Run.update_matrix_cdc` and `Run.get_domain_data` (reference run.py:499-594) are executed as they are - unbound from `Run`
and bound to a small stand-in object that carries only the attributes they read (n_domain, n_cluster, config.n_causal_mask,
domain_cnt_weight, device, per-domain loaders) - on a reference CDC(PLE) and CDC(MMoE) with seven domains.  Two consecutive
calls (the second one with groups installed, so matrix B gets its n_domain + n_cluster rows) and then one ordinary training
step, which shows that the Adam moments the probes left behind are the same.  NumPy's global RNG is seeded once; the loop draws
its domain multisets, shuffles and k-means seeds from it.  The affinity matrices are captured right before `update_group`
transforms them.  Output: tests/golden/cdc_probe_<base>.npz; tests/test_cdc_update_matrix.py replays them."""
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (installs the import shim, imports the reference models)
import run as ref_run  # noqa: E402  (the reference's run.py)

FIELD_DIMS = np.array([7, 5, 11, 7, 9, 6], dtype=np.int64)      # field 3 is the domain id: 7 domains
DOMAIN_IDX, ND, T, E = 3, 7, 3, 4
N_MASK, K_STEPS, N_BATCH = 4, 2, 3
WEIGHT = [0.25, 0.2, 0.15, 0.15, 0.1, 0.1, 0.05]


def domain_batches(rng):
    """N_BATCH batches per domain, 10 + d rows each, x[:, DOMAIN_IDX] == d."""
    out = []
    for d in range(ND):
        per = []
        for _ in range(N_BATCH):
            B = 10 + d
            x = np.stack([rng.integers(0, c, size=B) for c in FIELD_DIMS], axis=1).astype(np.int32)
            x[:, DOMAIN_IDX] = d
            y = (rng.random(B) < 0.35).astype(np.int16).reshape(B, 1)
            per.append((x, y))
        out.append(per)
    return out


def stand_in(batches):
    loaders = [[(torch.from_numpy(x), torch.from_numpy(y)) for x, y in per] for per in batches]
    me = types.SimpleNamespace(n_domain=ND, n_cluster=T, config=types.SimpleNamespace(n_causal_mask=N_MASK), domain_cnt_weight=WEIGHT,
                               device="cpu", train_data_loader=loaders, train_data_generator=[iter(ld) for ld in loaders],
                               domain2group_list=None)
    me.get_domain_data = types.MethodType(ref_run.Run.get_domain_data, me)
    me.update_matrix_cdc = types.MethodType(ref_run.Run.update_matrix_cdc, me)
    return me


def main():
    torch.set_num_threads(1)
    cfg = G.Cfg()
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    try:
        for base, ed, td in (("ple", ((16, 8), (8,)), (8, 4)), ("mmoe", (16, 8), (8, 4))):
            torch.manual_seed(2002)
            rng = np.random.default_rng(2002)
            batches = domain_batches(rng)
            c = G.CDC(FIELD_DIMS, E, T, ND, base, ed, td, DOMAIN_IDX, domain_cnt_weight=WEIGHT, n_causal_mask=N_MASK, device="cpu",
                      dropout=0.0, config=cfg, **G.L2)
            c.save_draw_matrix = lambda *a, **k: None
            out = {"meta": json.dumps(dict(nd=ND, T=T, E=E, n_mask=N_MASK, k=K_STEPS, weight=WEIGHT, domain_idx=DOMAIN_IDX))}
            for d, per in enumerate(batches):
                for i, (x, y) in enumerate(per):
                    out[f"data.{d}.{i}.x"], out[f"data.{d}.{i}.y"] = x, y
            for k, v in G.sd_np(c).items():
                out["sd0." + k] = v
            opt = torch.optim.Adam(c.parameters(), lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)
            crit = torch.nn.BCELoss()
            me = stand_in(batches)
            captured = {}
            inner = c.update_group

            def spy(*a, **k):
                captured["mask"], captured["A"], captured["B"] = (c.matrix_mask.numpy().copy(), c.matrix_A.numpy().copy(),
                                                                  c.matrix_B.numpy().copy())
                return inner(*a, **k)
            c.update_group = spy
            np.random.seed(77)
            c.train()                                                    # train_cdc: model.train() at the top of the epoch (run.py:598)
            for call in range(2):
                me.update_matrix_cdc(c, crit, opt, K_STEPS)
                for k in ("mask", "A", "B"):
                    out[f"call{call}.{k}"] = captured[k]
                out[f"call{call}.d2g"] = np.asarray(me.domain2group_list, dtype=np.int64)
                out[f"call{call}.s_groups"] = json.dumps([[int(v) for v in g] for g in c.s_group2domain_list])
                out[f"call{call}.training"] = np.asarray(c.training)
                for k, v in G.sd_np(c).items():
                    out[f"call{call}.sd." + k] = v
            # one ordinary step afterwards (run.py:635-640), model left in whatever mode the probing loop left it
            x, y = me.get_domain_data(2)
            p = c(x, mode="split", domain_i=2)
            loss = crit(p.squeeze(), y.squeeze().float()) + c.get_regularization_loss(device="cpu")
            c.zero_grad()
            loss.backward()
            opt.step()
            out["after.pred"] = p.detach().numpy().copy()
            out["after.loss"] = np.float32(loss.item())
            for k, v in G.sd_np(c).items():
                out["after.sd." + k] = v
            path = os.path.join(HERE, f"cdc_probe_{base}.npz")
            np.savez_compressed(path, **out)
            print(base, os.path.getsize(path) // 1024, "KiB", "d2g", out["call0.d2g"].tolist(), out["call1.d2g"].tolist(),
                  "training after:", bool(c.training))
    finally:
        os.chdir(cwd)


if __name__ == "__main__":
    main()
