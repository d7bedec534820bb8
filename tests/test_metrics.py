"""N4: Run.test's metrics on the device (cdcmdr_auc_logloss / cdcmdr_b200.metrics) against scikit-learn - the
functions the reference calls (run.py:682-705).  CPU: the host restatement of the entry point is pinned on roc_auc_score /
log_loss (ties, single-class domains, empty domains) and `evaluate_multi_domain` on the reference's own function run on the same
arrays; GPU: the kernel against the restatement (1e-12: both fp64 on the same formula) and against sklearn at 2 M samples."""
import numpy as np
import pytest
import torch
from sklearn.metrics import log_loss, roc_auc_score

import cdcmdr_b200 as cm
from oracle.host_abi import HostABI


@pytest.fixture
def emulator():
    old = cm._lib._LIB
    cm._lib.install(HostABI())
    yield
    cm._lib.install(old)


def _case(n, nd, seed, ties=True):
    rng = np.random.default_rng(seed)
    d = rng.integers(0, nd, size=n).astype(np.int64)
    y = (rng.random(n) < 0.2).astype(np.int16)
    p = (1 / (1 + np.exp(-(rng.standard_normal(n) + 1.5 * y)))).astype(np.float32)
    if ties:
        p = np.round(p, 2).astype(np.float32)                  # many tied predictions, some exactly 0 / 1 after rounding
    if nd > 3:
        y[d == 1] = 0                                          # a domain with one class only -> NaN
        d[d == 2] = 0                                          # an empty domain
    return p, y, d


def _sk(p, y, d, nd):
    out = np.full((nd, 2), np.nan)
    for k in range(nd):
        m = d == k
        if m.sum() and 0 < y[m].sum() < m.sum():
            out[k] = roc_auc_score(y[m], p[m]), log_loss(y[m], p[m])
    return out


@pytest.mark.parametrize("n,nd,ties", [(5000, 6, True), (777, 1, False), (3000, 30, True), (50, 4, False)])
def test_restatement_matches_sklearn(n, nd, ties, emulator):
    p, y, d = _case(n, nd, n + nd, ties)
    got = cm.metrics.auc_logloss(torch.from_numpy(p), torch.from_numpy(y), torch.from_numpy(d), nd).numpy()
    want = _sk(p, y, d, nd)
    assert np.array_equal(np.isnan(got[:, 0]), np.isnan(want[:, 0]))
    ok = ~np.isnan(want[:, 0])
    assert np.abs(got[ok, 0] - want[ok, 0]).max() <= 1e-12
    assert np.abs(got[ok, 1] - want[ok, 1]).max() <= 1e-6 * np.abs(want[ok, 1]).max()     # sklearn accumulates fp32 terms
    assert np.array_equal(got[:, 3], np.bincount(d, minlength=nd)[:nd])


def test_evaluate_multi_domain_matches_reference_function(emulator):
    """the dict Run.evaluate_multi_domain builds (run.py:690-711), from the reference's own function where it is importable"""
    import os
    import sys
    import types
    if not os.path.isdir("/root/reference"):
        pytest.skip("needs the reference checkout")
    for name in ("matplotlib", "matplotlib.pyplot", "wandb", "tqdm"):
        sys.modules.setdefault(name, types.ModuleType(name))
    ds, al, pp = types.ModuleType("dataset"), types.ModuleType("dataset.aliccp"), types.ModuleType("dataset.aliccp.preprocess_ali_ccp")
    pp.reduce_mem = lambda df: df
    sys.modules.setdefault("dataset", ds); sys.modules.setdefault("dataset.aliccp", al)
    sys.modules.setdefault("dataset.aliccp.preprocess_ali_ccp", pp)
    sys.path.insert(0, "/root/reference")
    try:
        from run import Run
    except Exception as exc:                                   # the reference's runner needs packages this image may lack
        pytest.skip(f"reference runner not importable: {exc!r}")
    nd = 7
    p, y, d = _case(4000, nd, 5)
    w = np.linspace(0.05, 0.25, nd); w = w / w.sum()

    class Stub:
        n_domain = nd
        domain_cnt_weight = w
    want = Run.evaluate_multi_domain(Stub(), y, p, d)
    got = cm.metrics.evaluate_multi_domain(torch.from_numpy(y), torch.from_numpy(p), torch.from_numpy(d), nd, w)
    assert set(got["domain_auc"]) == set(want["domain_auc"])
    for k in want["domain_auc"]:
        for key in ("domain_auc", "domain_loss"):
            a, b = got[key][k], want[key][k]
            assert (np.isnan(a) and np.isnan(b)) or abs(a - b) <= 1e-6 * max(1.0, abs(b)), (key, k, a, b)
    for key in ("mean_auc", "mean_loss"):
        a, b = got[key], want[key]
        assert (np.isnan(a) and np.isnan(b)) or abs(a - b) <= 1e-6 * max(1.0, abs(b))


@pytest.mark.gpu
@pytest.mark.parametrize("n,nd,ties", [(5000, 6, True), (1, 1, False), (200_000, 30, True), (2_000_000, 30, False)])
def test_device_metrics_match_restatement_and_sklearn(n, nd, ties):
    p, y, d = _case(n, nd, n % 1000 + nd, ties)
    got = cm.metrics.auc_logloss(torch.from_numpy(p).cuda(), torch.from_numpy(y).cuda(), torch.from_numpy(d).cuda(), nd).cpu().numpy()
    emu = HostABI()
    want = np.zeros((nd, 4))
    emu.auc_logloss(p.ctypes.data, y.ctypes.data, 0, d.ctypes.data, 1, n, nd, want.ctypes.data, None, 0)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.abs(got[ok] - want[ok]).max() <= 1e-9 * max(1.0, np.abs(want[ok]).max())
    if n >= 5000:
        sk = _sk(p, y, d, nd)
        k = ~np.isnan(sk[:, 0])
        assert np.abs(got[k, 0] - sk[k, 0]).max() <= 1e-10 and np.abs(got[k, 1] - sk[k, 1]).max() <= 2e-6 * np.abs(sk[k, 1]).max()
    # int32 domains, fp32 targets, and the single-set form
    g2 = cm.metrics.auc_logloss(torch.from_numpy(p).cuda(), torch.from_numpy(y.astype(np.float32)).cuda(),
                                torch.from_numpy(d.astype(np.int32)).cuda(), nd).cpu().numpy()
    assert np.array_equal(np.nan_to_num(g2, nan=-1), np.nan_to_num(got, nan=-1))
    tot = cm.metrics.auc_logloss(torch.from_numpy(p).cuda(), torch.from_numpy(y).cuda(), None, 1).cpu().numpy()[0]
    if 0 < y.sum() < n:
        assert abs(tot[0] - roc_auc_score(y, p)) <= 1e-10
