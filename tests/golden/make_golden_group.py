#!/usr/bin/env python
"""Golden fixtures for CDC.update_group (host-side clustering, reference model/cdc.py:121-341), produced by the UNMODIFIED
reference in the dev container:   python tests/golden/make_golden_group.py

For each case: seeded random affinity matrices are installed on a reference CDC model and update_group() is called three times
(first call: k-means on the causal distances; later calls: iterative / greedy regrouping).  NumPy's global RNG is re-seeded
before every call because upstream's KMeans is unseeded.  Inputs and every output list / matrix are dumped to
tests/golden/cdc_group_<case>.npz; tests/test_cdc_group.py replays them through cdcmdr_b200.cdc_group."""
import json
import zlib
import os
import sys
import tempfile
import types

import numpy as np
import torch

REF = os.environ.get("CDCMDR_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))
for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, REF)
from model.cdc import CDC  # noqa: E402


def cfg(affinity, p_method, old_w):
    class Cfg:
        use_atten = False; use_dcn = False; dataset_name = "synthetic"; mmoe_n_expert = 2
        p_weight = 0.1; p_weight_method = p_method; p_weight_exp_decay = 0.9; old_matrix_weight = old_w; affinity_func = affinity
    return Cfg()


CASES = {
    "minus_iter": dict(nd=12, nc=3, n_mask=20, affinity="minus", p_method="linear_decay", old_w=0.0, mode="iterative", metric="loss"),
    "minus_greedy_oldw": dict(nd=10, nc=3, n_mask=16, affinity="minus", p_method="exponential_decay", old_w=0.3, mode="greedy", metric="loss"),
    "divide_iter": dict(nd=9, nc=2, n_mask=14, affinity="divide", p_method="quadratic_decay", old_w=0.0, mode="iterative", metric="loss"),
}


def main():
    fd = np.array([5, 4, 6, 3], dtype=np.int64)
    for name, c in CASES.items():
        nd, nc = c["nd"], c["nc"]
        rng = np.random.default_rng(zlib.crc32(name.encode()) % 1000 + 7)
        w = rng.random(nd).astype(np.float32) + 0.2
        w /= w.sum()
        cwd = os.getcwd()
        with tempfile.TemporaryDirectory() as tmp:
            os.chdir(tmp)
            try:
                m = CDC(fd, 2, nc, nd, "mmoe", (4,), (4,), 3, domain_cnt_weight=w.tolist(), n_causal_mask=c["n_mask"],
                        use_metric=c["metric"], device="cpu", dropout=0.0, config=cfg(c["affinity"], c["p_method"], c["old_w"]))
                m.save_draw_matrix = lambda *a, **k: None
                dump = dict(meta=json.dumps(c), w=w)
                for call in range(3):
                    base = 0.5 + 0.05 * rng.standard_normal(nd).astype(np.float32)
                    A = (base[None, :] + 0.02 * rng.standard_normal((nd + 1, nd))).astype(np.float32)
                    B = (base[None, :] + 0.02 * rng.standard_normal((nd + nc, nd))).astype(np.float32)
                    M = (base[None, :] + 0.03 * rng.standard_normal((c["n_mask"], nd))).astype(np.float32)
                    m.matrix_A, m.matrix_B, m.matrix_mask = torch.from_numpy(A.copy()), torch.from_numpy(B.copy()), torch.from_numpy(M.copy())
                    np.random.seed(100 + call)
                    d2g = m.update_group(mode=c["mode"])
                    dump[f"in{call}.A"], dump[f"in{call}.B"], dump[f"in{call}.M"] = A, B, M
                    dump[f"out{call}.d2g"] = np.asarray(d2g, dtype=np.int64)
                    dump[f"out{call}.s_groups"] = json.dumps([[int(v) for v in g] for g in m.s_group2domain_list])
                    dump[f"out{call}.t_groups"] = json.dumps([[int(v) for v in g] for g in m.t_group2domain_list])
                    dump[f"out{call}.A"] = m.matrix_A.numpy().copy()
                    dump[f"out{call}.B"] = m.matrix_B.numpy().copy()
                    dump[f"out{call}.mask"] = m.matrix_mask.numpy().copy()
                    dump[f"out{call}.causal"] = m.matrix_causal.numpy().copy()
                    dump[f"out{call}.p_weight"] = np.float64(m.p_weight)
            finally:
                os.chdir(cwd)
        np.savez_compressed(os.path.join(OUT, f"cdc_group_{name}.npz"), **dump)
        print(name, "ok")


if __name__ == "__main__":
    main()
