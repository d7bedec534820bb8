"""CDC.update_group (host-side causal domain clustering, reference model/cdc.py:121-341) replayed against fixtures
produced by the unmodified reference (tests/golden/make_golden_group.py): three consecutive calls per case - k-means
initialisation, then iterative / greedy regrouping with source-domain growth, old-matrix blending and p_weight decay."""
import json
import os

import numpy as np
import pytest
import torch

import cdcmdr_b200 as cm
from oracle.host_abi import HostABI
from tests.conftest import GOLDEN

CASES = ["minus_iter", "minus_greedy_oldw", "divide_iter"]


@pytest.fixture(autouse=True)
def emulator():
    old = cm._lib._LIB
    cm._lib.install(HostABI())
    yield
    cm._lib.install(old)


@pytest.mark.parametrize("name", CASES)
def test_update_group_matches_reference(name):
    z = np.load(os.path.join(GOLDEN, f"cdc_group_{name}.npz"))
    c = json.loads(str(z["meta"]))
    nd, nc = c["nd"], c["nc"]

    class Cfg:
        use_atten = False; use_dcn = False; dataset_name = "synthetic"; mmoe_n_expert = 2
        p_weight = 0.1; p_weight_method = c["p_method"]; p_weight_exp_decay = 0.9; old_matrix_weight = c["old_w"]
        affinity_func = c["affinity"]
    fd = np.array([5, 4, 6, 3], dtype=np.int64)
    m = cm.CDC(fd, 2, nc, nd, "mmoe", (4,), (4,), 3, domain_cnt_weight=z["w"].tolist(), n_causal_mask=c["n_mask"],
               use_metric=c["metric"], device="cpu", dropout=0.0, config=Cfg())
    for call in range(3):
        m.matrix_A = torch.from_numpy(z[f"in{call}.A"].copy())
        m.matrix_B = torch.from_numpy(z[f"in{call}.B"].copy())
        m.matrix_mask = torch.from_numpy(z[f"in{call}.M"].copy())
        np.random.seed(100 + call)                       # upstream's KMeans is unseeded
        d2g = m.update_group(mode=c["mode"])
        assert d2g == z[f"out{call}.d2g"].tolist(), (call, d2g, z[f"out{call}.d2g"].tolist())
        assert m.domain2group.tolist() == d2g and m.domain2group.dtype == torch.int64
        assert m.s_group2domain_list == json.loads(str(z[f"out{call}.s_groups"])), call
        assert m.t_group2domain_list == json.loads(str(z[f"out{call}.t_groups"])), call
        np.testing.assert_allclose(m.matrix_A.numpy(), z[f"out{call}.A"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(m.matrix_B.numpy(), z[f"out{call}.B"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(m.matrix_mask.numpy(), z[f"out{call}.mask"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(m.matrix_causal.numpy(), z[f"out{call}.causal"], rtol=1e-5, atol=1e-6)
        assert abs(m.p_weight - float(z[f"out{call}.p_weight"])) < 1e-12
        assert m.call_update_group == call + 1


def test_causal_kernel_matches_oracle():
    from oracle import cdcmdr_oracle as O
    from cdcmdr_b200_pkg.cdc_group import calc_causal_matrix
    rng = np.random.default_rng(0)
    X = rng.standard_normal((9, 14))
    np.testing.assert_allclose(calc_causal_matrix(X), O.calc_causal_matrix(X), rtol=1e-12)
