"""GPU timing of the CDC affinity-matrix probing loop at the C4 shape (not a test, not the bench contract):

    python tools/bench_probe.py [--batch 65536] [--masks 50] [--k 1]

CDC(base=PLE), 30 domains -> 4 clusters, 23 fields x embed 16, vocab 1 M, bf16 path.  One `CDC.update_matrix_cdc` call =
snapshot + (masks + 30 + 31..34) probes of {k fused training steps on a domain (multi)set, ONE batched evaluation of all 30
domains' batches (30 x batch rows) with per-domain BCE means on the device, restore} + update_group on the host.  The first call
is the warm-up (workspaces for every batch size are allocated then), the second is timed with a device synchronise on both sides.
Also timed, on the same weights and batches: one affinity-matrix row the reference's way (30 forwards of one domain each, each
followed by a host BCE read - run.py:550-558) against the batched probe.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cdcmdr_b200 as cm  # noqa: E402

F, E, T, ND, DOM = 23, 16, 4, 30, 10
DEV = os.environ.get("CDCMDR_DEVICE", "cuda")


def sync():
    if DEV == "cuda":
        torch.cuda.synchronize()


class Cfg:
    use_atten = False; use_dcn = False; ple_n_expert_specific = 2; ple_n_expert_shared = 2
    cdcmdr_precision = "bf16"
    p_weight = 0.1; p_weight_method = "linear_decay"; p_weight_exp_decay = 0.9; old_matrix_weight = 0.0; affinity_func = "minus"


class Provider:
    """run.py:499-526 over device-resident batches: an int -> that domain's next batch, a list -> shuffled, concatenated."""

    def __init__(self, fd, B, n_per_domain, seed):
        rng = np.random.default_rng(seed)
        self.loaders = []
        for d in range(ND):
            per = []
            for _ in range(n_per_domain):
                x = np.stack([np.minimum(rng.zipf(1.05, size=B) - 1, c - 1) for c in fd], axis=1).astype(np.int32)
                x[:, DOM] = d
                y = (rng.random(B) < 0.05 + 0.002 * d).astype(np.int16).reshape(B, 1)
                per.append((torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV)))
            self.loaders.append(per)
        self.pos = [0] * ND
        self.rows = 0

    def __call__(self, d):
        if isinstance(d, (int, np.integer)):
            b = self.loaders[d][self.pos[d] % len(self.loaders[d])]
            self.pos[d] += 1
            self.rows += b[0].shape[0]
            return b
        np.random.shuffle(d)
        got = [self(int(i)) for i in d]
        return torch.cat([g[0] for g in got], dim=0), torch.cat([g[1] for g in got], dim=0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--masks", type=int, default=50)
    ap.add_argument("--k", type=int, default=1)
    ap.add_argument("--dropout", type=float, default=0.2)
    ap.add_argument("--vocab", type=int, default=45_454, help="rows per field")
    a = ap.parse_args()
    fd = np.full(F, a.vocab, dtype=np.int64); fd[DOM] = ND
    torch.manual_seed(2000)
    w = np.arange(ND, 0, -1, dtype=np.float64) ** 1.2
    w = (w / w.sum()).tolist()
    m = cm.CDC(fd, E, T, ND, "ple", ((256, 128), (64,)), (64, 32), DOM, domain_cnt_weight=w, n_causal_mask=a.masks, dropout=a.dropout,
               config=Cfg(), l2_reg_embedding=1e-5, l2_reg_linear=1e-5, l2_reg_dnn=1e-5).to(DEV).train()
    opt = cm.Adam(m.parameters(), lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)
    get = Provider(fd, a.batch, 2, 11)
    np.random.seed(2000)
    lib = m.base_model_instance._rt.ops.lib
    host = {}
    inner = m.update_group

    def timed_group(*args, **kw):
        sync()
        t0 = time.perf_counter()
        try:
            return inner(*args, **kw)
        finally:
            host["update_group_s"] = time.perf_counter() - t0
    m.update_group = timed_group
    for _ in range(3):                                                   # warm-up steps (run.py:609-627)
        m.train_step(*get(0), opt, mode="warmup")
    res = []
    for call in range(2):
        get.rows = 0
        lib.launch_count_reset()
        sync()
        t0 = time.perf_counter()
        d2g = m.update_matrix_cdc(get, opt, a.k)
        sync()
        dt = time.perf_counter() - t0
        n_probe = a.masks + 1 + ND + (ND + T if call else ND + 1)
        res.append(dict(seconds=dt, probes=n_probe, rows=get.rows, launches=int(lib.launch_count()), update_group_s=host["update_group_s"],
                        groups=sorted(np.bincount(d2g, minlength=T).tolist())))
    # one matrix row, the reference's way vs the batched probe
    m.eval()
    batches = [get(d) for d in range(ND)]

    def per_domain():
        row = np.zeros(ND)
        with torch.no_grad():
            for d, (x, y) in enumerate(batches):
                p = m(x, mode="split", domain_i=d)
                row[d] = float(m.get_matrix_metric(p.squeeze(), y.squeeze().float()))
        return row

    def batched():
        return m.probe_all_domains(batches).cpu().numpy()
    t = {}
    for name, fn in (("per_domain", per_domain), ("batched", batched)):
        fn()
        sync()
        t0 = time.perf_counter()
        for _ in range(5):
            row = fn()
        sync()
        t[name] = (time.perf_counter() - t0) / 5
        t[name + "_row"] = row
    err = float(np.abs(t["per_domain_row"] - t["batched_row"]).max())
    timed = res[1]
    print(json.dumps(dict(what="CDC.update_matrix_cdc at C4 (CDC-PLE, 30 domains, bf16)", batch_per_domain=a.batch, k=a.k, masks=a.masks,
                          warm_call=res[0], timed_call=timed, probes_per_s=timed["probes"] / timed["seconds"],
                          rows_per_s=timed["rows"] / timed["seconds"],
                          matrix_row_ms=dict(per_domain=1e3 * t["per_domain"], batched=1e3 * t["batched"], max_abs_diff=err),
                          mem_gb=(torch.cuda.max_memory_allocated() / 2 ** 30) if DEV == "cuda" else 0.0)))


if __name__ == "__main__":
    main()
