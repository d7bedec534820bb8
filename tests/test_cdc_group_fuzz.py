"""Differential fuzzing of `cdc_group.Grouping.update` against the reference's own `CDC.update_group` (model/cdc.py:121-341) run
live here: random sizes (6..15 domains, 2..4 clusters), both affinity functions, all three p_weight schedules, old-matrix
blending on / off, iterative / greedy regrouping, loss / auc orientation, three consecutive calls each (k-means initialisation,
then two regroupings).  Needs /root/reference (skipped elsewhere; the committed fixtures of tests/test_cdc_group.py travel).

Why it exists: the regrouping compares sums over the same domains taken in different orders, so its decisions depend on the last
bit of float32 reductions and on torch's NaN rules for argmin / min - three fixtures did not reach either case; this sweep did
(3 of 480 calls with NumPy reductions, and the NaN case through the probing-loop fixtures) and pins the torch-tensor arithmetic
that replaced them."""
import contextlib
import io
import os
import sys
import tempfile

import numpy as np
import pytest
import torch

import cdcmdr_b200 as cm

REF = os.environ.get("CDCMDR_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "model", "cdc.py")), reason="needs the reference checkout")
FD = np.array([5, 4, 6, 3], dtype=np.int64)


def _case(seed):
    rng = np.random.default_rng(seed)
    nd, nc = int(rng.integers(6, 16)), int(rng.integers(2, 5))
    return rng, dict(nd=nd, nc=nc, n_mask=int(rng.integers(nd, 30)), aff=["minus", "divide"][seed % 2],
                     pm=["linear_decay", "quadratic_decay", "exponential_decay"][seed % 3], oldw=[0.0, 0.3][(seed // 2) % 2],
                     mode=["iterative", "greedy"][(seed // 3) % 2], metric=["loss", "auc"][(seed // 5) % 2])


@pytest.mark.parametrize("block", range(3))
def test_grouping_matches_live_reference(block, monkeypatch):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden as G                     # import shim + the reference's CDC
    monkeypatch.chdir(tempfile.mkdtemp())
    for seed in range(20 * block + 30, 20 * block + 50):          # seeds 30..89 (33, 43 and 49 tripped the NumPy-sum version)
        rng, c = _case(seed)
        nd, nc = c["nd"], c["nc"]

        class Cfg:
            use_atten = False; use_dcn = False; dataset_name = "synthetic"; mmoe_n_expert = 2
            p_weight = 0.1; p_weight_method = c["pm"]; p_weight_exp_decay = 0.9; old_matrix_weight = c["oldw"]; affinity_func = c["aff"]
        w = rng.random(nd).astype(np.float32) + 0.2
        w /= w.sum()
        with contextlib.redirect_stdout(io.StringIO()):
            ref = G.CDC(FD, 2, nc, nd, "mmoe", (4,), (4,), 3, domain_cnt_weight=w.tolist(), n_causal_mask=c["n_mask"],
                        use_metric=c["metric"], device="cpu", dropout=0.0, config=Cfg())
        ref.save_draw_matrix = lambda *a, **k: None
        mine = cm.cdc_group.Grouping(nd, nc, w.tolist(), Cfg(), c["metric"])
        for call in range(3):
            base = 0.5 + 0.05 * rng.standard_normal(nd).astype(np.float32)
            A = (base[None, :] + 0.02 * rng.standard_normal((nd + 1, nd))).astype(np.float32)
            B = (base[None, :] + 0.02 * rng.standard_normal((nd + nc, nd))).astype(np.float32)
            M = (base[None, :] + 0.03 * rng.standard_normal((c["n_mask"], nd))).astype(np.float32)
            ref.matrix_A, ref.matrix_B, ref.matrix_mask = torch.from_numpy(A.copy()), torch.from_numpy(B.copy()), torch.from_numpy(M.copy())
            outcome = []
            for run in ("ref", "mine"):
                np.random.seed(100 + call)               # upstream's KMeans is unseeded: NumPy's global RNG
                try:
                    if run == "ref":
                        with contextlib.redirect_stdout(io.StringIO()):
                            d2g = ref.update_group(mode=c["mode"])
                        outcome.append(([int(v) for v in d2g], [[int(v) for v in g] for g in ref.s_group2domain_list],
                                        [[int(v) for v in g] for g in ref.t_group2domain_list]))
                    else:
                        out = mine.update(A.copy(), B.copy(), M.copy(), mode=c["mode"])
                        outcome.append(([int(v) for v in mine.domain2group_list], [[int(v) for v in g] for g in mine.s_group2domain_list],
                                        [[int(v) for v in g] for g in mine.t_group2domain_list]))
                except ValueError as e:                  # upstream raises when the iterative regrouping stalls (cdc.py:203-206)
                    outcome.append(("ValueError", str(e)))
            assert outcome[0] == outcome[1], (seed, call, c, outcome)
            if outcome[0][0] == "ValueError":
                break
            for k, t in (("A", ref.matrix_A), ("B", ref.matrix_B), ("mask", ref.matrix_mask), ("causal", ref.matrix_causal)):
                assert np.allclose(out[k], t.numpy(), rtol=0, atol=1e-6, equal_nan=True), (seed, call, k)
