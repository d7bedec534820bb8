"""CDC steps whose domain - hence the selected tower - changes every step (reference run.py:635-640 with a per-domain batch
sequence), against fixtures produced by the unmodified reference (tests/golden/make_golden_alternate.py).  In the reference the
towers a step does not select still receive ZERO gradients (select / cat backward), not None, so Adam keeps decaying their
moments and they coast; two identical steps cannot show that, five steps over towers [2, 0, 1, 2, 2] do."""
import os

import numpy as np
import pytest
import torch

import cdcmdr_b200 as cm
from oracle.host_abi import HostABI
from tests.conftest import GOLDEN
from tests.golden_cases import E, FIELD_DIMS, L2
from tests.test_oracle_golden import bias_before_bn
from tests.util import Cfg, DOMAIN_IDX

BASES = {"ple": (((16, 8), (8,)), (8, 4)), "mmoe": ((16, 8), (8, 4)), "star": (None, (16, 8))}


def _run(base, device, path):
    z = np.load(os.path.join(GOLDEN, f"cdc_{base}_alternate.npz"))
    ed, td = BASES[base]
    d2g, domains = z["d2g"].tolist(), z["domains"].tolist()
    m = cm.CDC(FIELD_DIMS, E, 3, len(d2g), base, ed, td, DOMAIN_IDX, domain_cnt_weight=[0.4, 0.3, 0.2, 0.1], dropout=0.0,
               config=Cfg(), **L2)
    m.set_groups(d2g)
    m.load_state_dict({k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd0.")}, strict=True)
    m = m.to(device).train()
    opt = cm.Adam(m.parameters(), lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)
    opt.attach(m.base_model_instance)
    crit = torch.nn.BCELoss()
    for s, dom in enumerate(domains):
        x, y = torch.from_numpy(z[f"in{s}.x"]).to(device), torch.from_numpy(z[f"in{s}.y"]).to(device)
        if path == "fused":
            out = m.train_step(x, y, opt, mode="split", domain_i=dom)
            _, bce, reg = m.step_losses(out)
            pred = out["psel"].cpu().numpy()
        else:
            p = m(x, mode="split", domain_i=dom)
            bce_t = crit(p, y.squeeze().float())
            reg_t = m.get_regularization_loss(device=device)
            m.zero_grad()
            (bce_t + reg_t).backward()
            opt.step()
            bce, reg, pred = float(bce_t.detach()), float(reg_t.detach()), p.detach().cpu().numpy()
        assert np.abs(pred - z[f"step{s}.pred"]).max() <= 1e-4 + 2e-3 * s, (s, float(np.abs(pred - z[f"step{s}.pred"]).max()))
        assert abs(bce - float(z[f"step{s}.bce"])) <= 1e-4 + 1e-3 * s and abs(reg - float(z[f"step{s}.reg"])) <= 1e-4 * float(z[f"step{s}.reg"])
        cur = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
        kind = base
        for k, v in cur.items():
            ref = z[f"sd{s + 1}." + k]
            if k.endswith("num_batches_tracked"):
                assert int(v) == int(ref), (s, k)
                continue
            kk = k[len("base_model_instance."):]
            tol = 2.1e-3 * (s + 1) if (bias_before_bn(kind, kk) or kk.endswith("running_mean")) else 2e-5 * (s + 1)
            assert np.abs(v - ref).max() <= tol, (s, k, float(np.abs(v - ref).max()))


@pytest.fixture
def emulator():
    old = cm._lib._LIB
    cm._lib.install(HostABI())
    yield
    cm._lib.install(old)


@pytest.mark.parametrize("path", ["fused", "autograd"])
@pytest.mark.parametrize("base", sorted(BASES))
def test_alternating_towers_host_logic(base, path, emulator):
    _run(base, "cpu", path)


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["fused", "autograd"])
@pytest.mark.parametrize("base", sorted(BASES))
def test_alternating_towers_gpu(base, path):
    _run(base, "cuda", path)
