"""CDC affinity probe (reference run.py:551-560 `cdc_test_all_domain`): `CDC.probe_all_domains` - one batched
evaluation over the concatenated per-domain batches + per-domain BCE means on the device - against the reference's own
procedure restated with the package's public API: one `model(X_d, mode='split', domain_i=d)` per domain in eval mode followed
by `get_matrix_metric` (torch BCELoss).  Eval-mode BatchNorm uses running statistics, so batching must not change a bit of the
predictions; the means agree to fp32 summation order."""
import numpy as np
import pytest
import torch

import cdcmdr_b200 as cm
from oracle.host_abi import HostABI
from tests.golden_cases import E, FIELD_DIMS, L2
from tests.util import Cfg, DOMAIN_IDX

BASES = {"ple": (((16, 8), (8,)), (8, 4)), "mmoe": ((16, 8), (8, 4))}


def _run(base, device, sizes):
    ed, td = BASES[base]
    n_domain = int(FIELD_DIMS[DOMAIN_IDX])
    d2g = [d % 3 for d in range(n_domain)]
    torch.manual_seed(5)
    m = cm.CDC(FIELD_DIMS, E, 3, n_domain, base, ed, td, DOMAIN_IDX, domain_cnt_weight=[1.0 / n_domain] * n_domain,
               dropout=0.2 if device == "cuda" else 0.0,        # the emulator has no dropout; on the GPU eval mode must switch it off
               config=Cfg(), **L2)
    m.set_groups(d2g)
    m = m.to(device).train()
    rng = np.random.default_rng(17)
    batches = []
    for d in range(n_domain):
        B = sizes[d % len(sizes)]
        x = np.stack([rng.integers(0, fd, size=B) for fd in FIELD_DIMS], axis=1).astype(np.int32)
        x[:, DOMAIN_IDX] = d
        y = (rng.random(B) < 0.3).astype(np.int16)
        batches.append((torch.from_numpy(x).to(device), torch.from_numpy(y).to(device)))
    # a few training steps first so that the BatchNorm running statistics and the weights are not at their initial values
    opt = cm.Adam(m.parameters(), lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)
    for d in (0, 1, 2):
        if batches[d][0].shape[0] > 1:
            m.train_step(batches[d][0], batches[d][1], opt, mode="split", domain_i=d)
    got = m.probe_all_domains(batches)
    assert m.training                                           # the probe restores the mode it found
    # capped evaluations (consecutive domains grouped up to max_rows rows; a single batch may exceed the cap): same numbers
    for cap in (1, sum(sizes), 3 * max(sizes)):
        chunked = m.probe_all_domains(batches, max_rows=cap)
        assert torch.allclose(torch.nan_to_num(chunked, nan=-1.0), torch.nan_to_num(got, nan=-1.0), rtol=0, atol=1e-6), cap
    m.eval()
    want = []
    with torch.no_grad():
        for d, (x, y) in enumerate(batches):
            if x.shape[0] == 0:
                want.append(float("nan"))
                continue
            pred = m(x, mode="split", domain_i=d)
            want.append(float(m.get_matrix_metric(pred.reshape(-1), y.reshape(-1).float())))
    got = got.cpu().numpy()
    want = np.asarray(want, dtype=np.float32)
    assert got.shape == want.shape
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.abs(got[ok] - want[ok]).max() <= 2e-6 * max(1.0, float(np.abs(want[ok]).max())), (got, want)


@pytest.fixture
def emulator():
    old = cm._lib._LIB
    cm._lib.install(HostABI())
    yield
    cm._lib.install(old)


@pytest.mark.parametrize("sizes", [(37, 5, 64), (30, 0, 17)])
@pytest.mark.parametrize("base", sorted(BASES))
def test_probe_all_domains_host_logic(base, sizes, emulator):
    _run(base, "cpu", sizes=sizes)


@pytest.mark.gpu
@pytest.mark.parametrize("sizes", [(37, 5, 64), (4096, 1000, 333), (300, 0, 17)])
@pytest.mark.parametrize("base", sorted(BASES))
def test_probe_all_domains_gpu(base, sizes):
    _run(base, "cuda", sizes)


def test_probe_with_auc_metric_runs_per_domain_on_the_host(emulator):
    """use_metric='auc' (cdc.py:116-119): scikit-learn per domain on the host, as upstream - no batched device reduction"""
    from sklearn.metrics import roc_auc_score
    m = cm.CDC(FIELD_DIMS, E, 3, 4, "ple", ((16, 8), (8,)), (8, 4), DOMAIN_IDX, use_metric="auc", dropout=0.0, config=Cfg(), **L2)
    m.set_groups([0, 1, 2, 1])
    rng = np.random.default_rng(0)
    batches = []
    for d in range(4):
        x = np.stack([rng.integers(0, c, size=30) for c in FIELD_DIMS], axis=1).astype(np.int32)
        x[:, DOMAIN_IDX] = d
        batches.append((torch.from_numpy(x), torch.from_numpy((rng.random(30) < 0.4).astype(np.int16))))
    m.train()
    row = m.probe_all_domains(batches).numpy()
    assert m.training
    m.eval()
    with torch.no_grad():
        want = [roc_auc_score(y.numpy(), m(x, mode="split", domain_i=d).numpy().reshape(-1)) for d, (x, y) in enumerate(batches)]
    assert np.allclose(row, want, atol=1e-6)
