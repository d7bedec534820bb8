"""Model-level parity on the GPU through the package's public API (which calls the C-ABI of libcdcmdr.so):
  * every golden fixture produced by the unmodified reference (tests/golden/*.npz), through the fused train step and
    through the drop-in autograd path (`loss.backward(); optimizer.step()`);
  * larger seeded random cases against the CPU oracle (oracle/cdcmdr_oracle.py);
  * size-independent properties at BASELINE-sized batches (eval determinism, selection modes agree, loss decreases).
fp32 path: logits / gradients within 1e-4 relative (north_star)."""
import numpy as np
import pytest
import torch

import cdcmdr_b200 as cm
from oracle import cdcmdr_oracle as O
from tests.golden_cases import CASES
from tests.util import build_model, run_golden_case

pytestmark = pytest.mark.gpu
SUPPORTED = sorted(n for n in CASES if build_model(n, probe=True))


@pytest.mark.parametrize("name", SUPPORTED)
def test_golden_fused(name):
    run_golden_case(name, device="cuda", path="fused")


@pytest.mark.parametrize("name", SUPPORTED)
def test_golden_autograd(name):
    run_golden_case(name, device="cuda", path="autograd")


def _f64(sd):
    return {k: (v.astype(np.float64) if v.dtype == np.float32 else v.copy()) for k, v in sd.items()}


def _rand_ple(seed, B, F, E, T, vocab, dims, tower, device):
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed)
    fd = np.full(F, vocab, dtype=np.int64)
    model = cm.PLE(fd, E, T, 2, 2, dims, tower, dropout=0.0, l2_reg_embedding=1e-5, l2_reg_linear=1e-5, l2_reg_dnn=1e-5)
    x = rng.integers(0, vocab, size=(B, F)).astype(np.int32)
    y = (rng.random(B) < 0.2).astype(np.int16)
    g = rng.integers(0, T, size=B).astype(np.int64)
    return model, fd, x, y, g


@pytest.mark.parametrize("B,F,E,T", [(2048, 16, 16, 4), (1000, 23, 16, 3), (513, 5, 8, 2)])
def test_ple_gradients_and_steps_match_oracle(B, F, E, T):
    """Seeded random PLE at larger sizes than the fixtures, against the CPU oracle:
    (1) drop-in autograd path, step 0: predictions, loss and EVERY gradient (incl. the dense embedding gradient) within
        1e-4 of each tensor's scale;
    (2) fused path, 3 steps: step-0 predictions within 1e-4; later steps only statistically - Adam's first steps move
        every entry by ~lr*sign(g), so the few entries whose gradient is at rounding-noise level flip by up to 2*lr per
        step in a summation-order dependent way (in the reference as much as here)."""
    dims, tower = ((64, 32), (16,)), (16, 8)
    model, fd, x, y, g = _rand_ple(1, B, F, E, T, 500, dims, tower, "cuda")
    sd = {k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}
    om = O.PLE(fd, E, T, 2, 2, dims, tower, l2_reg_embedding=1e-5, l2_reg_linear=1e-5, l2_reg_dnn=1e-5)
    model = model.to("cuda").train()
    xt, yt, gt = (torch.from_numpy(a).cuda() for a in (x, y, g))
    # (1) gradients through loss.backward()
    # the oracle runs in float64 here: in float32 numpy a tower unit whose BatchNorm input is (nearly) constant over the
    # batch gets a rounding-noise xhat and a coin-flip ReLU mask, which moves its gradients by percents; float64 pins the
    # exact value of the same algorithm (the unmodified reference in fp32 AND in fp64 agrees with it to ~3e-6)
    r = O.train_step(om, _f64(sd), None, x, y, "gather", group=g)
    pred = model(xt)
    loss = torch.nn.BCELoss()(pred.gather(1, gt[:, None]).squeeze(1), yt.float()) + model.get_regularization_loss(device="cuda")
    model.zero_grad()
    loss.backward()
    assert float(np.abs(pred.detach().cpu().numpy() - r["pred"]).max()) <= 1e-4
    assert abs(float(loss) - float(r["loss"])) <= 1e-4 * abs(float(r["loss"]))
    named = dict(model.named_parameters())
    assert {k for k, p in named.items() if p.grad is not None} == set(r["grads"])
    for k, go in r["grads"].items():
        gm = named[k].grad.detach().cpu().numpy().reshape(go.shape)
        scale = float(np.abs(go).max())
        noise = 5e-5 if (k.startswith("towers") and k.endswith(".bias") and not k.endswith("layers.8.bias")) else 0.0
        err = float(np.abs(gm - go).max())
        assert err <= 1e-4 * scale + 1e-9 + noise, (k, err, scale)
    # the regulariser state must not have been disturbed: fused steps start from the same weights
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    opt = cm.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)
    oopt = O.Adam()
    sd = _f64(sd)
    for s in range(3):
        r = O.train_step(om, sd, oopt, x, y, "gather", group=g)
        out = model.train_step(xt, yt, opt, mode="gather", sel=gt)
        loss, bce, reg = model.step_losses(out)
        perr = float(np.abs(out["pred"].cpu().numpy() - r["pred"]).max())
        assert perr <= (1e-4 if s == 0 else 3e-2), (s, perr)
        assert abs(bce - float(r["bce"])) <= (1e-4 if s == 0 else 5e-3) * abs(float(r["bce"])) + 1e-6
        assert abs(reg - float(r["reg"])) <= 1e-4 * abs(float(r["reg"])) + 1e-6
        cur = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
        for k, v in sd.items():
            if k.endswith("num_batches_tracked"):
                assert int(cur[k]) == int(v)
                continue
            d = np.abs(cur[k].astype(np.float64) - v)
            assert d.max() <= 2.05e-3 * (s + 1) + 1e-5, (s, k, d.max())
            if d.size >= 64 and not k.endswith(".bias"):
                assert (d > 2e-5 * (s + 1)).mean() <= 0.03, (s, k, float((d > 2e-5 * (s + 1)).mean()))


def test_eval_forward_is_bit_repeatable_and_matches_train_free_path():
    model, fd, x, y, g = _rand_ple(2, 4096, 16, 16, 4, 1000, ((64, 32), (16,)), (16, 8), "cuda")
    model = model.to("cuda").eval()
    xt = torch.from_numpy(x).cuda()
    with torch.no_grad():
        a = model(xt).clone()
        b = model(xt).clone()
    assert torch.equal(a, b)
    assert a.shape == (4096, 4) and bool(((a > 0) & (a < 1)).all())


def test_cdc_selection_modes_agree_with_full_prediction():
    class Cfg:
        use_atten = False; use_dcn = False; ple_n_expert_specific = 2; ple_n_expert_shared = 2; mmoe_n_expert = 4
    torch.manual_seed(3)
    rng = np.random.default_rng(3)
    F, E, T, nd, B = 23, 16, 4, 30, 8192
    fd = np.full(F, 100, dtype=np.int64); fd[10] = nd
    m = cm.CDC(fd, E, T, nd, "ple", ((64, 32), (16,)), (16, 8), 10, dropout=0.0, config=Cfg()).to("cuda").eval()
    m.set_groups(rng.integers(0, T, size=nd).tolist())
    x = np.stack([rng.integers(0, d, size=B) for d in fd], axis=1).astype(np.int32)
    xt = torch.from_numpy(x).cuda()
    with torch.no_grad():
        full = m.base_model_instance(xt)
        warm = m(xt, mode="warmup")
        split = m(xt, mode="split")
        dom = m(xt, mode="split", domain_i=7)
    d2g = np.array(m.domain2group_list)
    fulln = full.cpu().numpy()
    assert np.array_equal(split.cpu().numpy()[:, 0], fulln[np.arange(B), d2g[x[:, 10]]])     # routing: bit-exact
    assert np.array_equal(dom.cpu().numpy(), fulln[:, d2g[7]])
    np.testing.assert_allclose(warm.cpu().numpy(), fulln.mean(1), rtol=1e-6)


def test_full_size_step_runs_and_learns():
    """C4-sized batch (B=65536, F=23, E=16, CDC-PLE stock dims): loss is finite and decreases over a few steps."""
    class Cfg:
        use_atten = False; use_dcn = False; ple_n_expert_specific = 2; ple_n_expert_shared = 2
    torch.manual_seed(4)
    rng = np.random.default_rng(4)
    F, E, T, nd, B = 23, 16, 4, 30, 65536
    fd = np.full(F, 20000, dtype=np.int64); fd[10] = nd
    m = cm.CDC(fd, E, T, nd, "ple", ((256, 128), (64,)), (64, 32), 10, dropout=0.0, config=Cfg(),
               l2_reg_embedding=1e-7, l2_reg_linear=1e-7, l2_reg_dnn=1e-7).to("cuda").train()
    m.set_groups([d % T for d in range(nd)])
    opt = cm.Adam(m.parameters(), lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)
    x = np.stack([rng.integers(0, d, size=B) for d in fd], axis=1).astype(np.int32)
    x[:, 10] = 5
    y = ((x[:, 0] % 7 == 0) | (rng.random(B) < 0.02)).astype(np.int16)
    xt, yt = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    losses = []
    for _ in range(8):
        out = m.train_step(xt, yt, opt, mode="split", domain_i=5)
        losses.append(m.step_losses(out)[1])
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses


def _auc(p, y):
    order = np.argsort(p, kind="stable")
    ranks = np.empty(len(p)); ranks[order] = np.arange(1, len(p) + 1)
    pos = y > 0
    return float((ranks[pos].sum() - pos.sum() * (pos.sum() + 1) / 2) / (pos.sum() * (~pos).sum()))


@pytest.mark.parametrize("kind", ["ple", "mmoe"])
def test_bf16_tensor_core_path_matches_fp32_oracle(kind):
    """bf16 tcgen05 path vs the fp32 CPU oracle on identical inputs / weights: logits within 2e-2 of the logit scale
    (north_star), BCE within 2e-2 relative, regulariser (fp32 master weights) within 1e-4."""
    class Cfg:
        use_atten = False; use_dcn = False; cdcmdr_precision = "bf16"
    torch.manual_seed(11)
    rng = np.random.default_rng(11)
    F, E, T, B = 16, 16, 4, 4096
    fd = np.full(F, 300, dtype=np.int64)
    l2 = dict(l2_reg_embedding=1e-5, l2_reg_linear=1e-5, l2_reg_dnn=1e-5)
    if kind == "ple":
        dims, tower = ((128, 64), (32,)), (32, 16)
        model = cm.PLE(fd, E, T, 2, 2, dims, tower, dropout=0.0, config=Cfg(), **l2)
        om = O.PLE(fd, E, T, 2, 2, dims, tower, **l2)
    else:
        dims, tower = (128, 64, 32), (32, 16)
        model = cm.MMoE(fd, E, T, 4, dims, tower, dropout=0.0, config=Cfg(), **l2)
        om = O.MMoE(fd, E, T, 4, dims, tower, **l2)
    sd = {k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}
    model = model.to("cuda").train()
    x = rng.integers(0, 300, size=(B, F)).astype(np.int32)
    y = (rng.random(B) < 0.2).astype(np.int16)
    g = rng.integers(0, T, size=B).astype(np.int64)
    r = O.train_step(om, sd, O.Adam(), x, y, "gather", group=g)
    opt = cm.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)
    out = model.train_step(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), opt, mode="gather", sel=torch.from_numpy(g).cuda())
    _, bce, reg = model.step_losses(out)
    logit = lambda p: np.log(p / (1 - p))                                        # noqa: E731
    lo, lr_ = logit(out["pred"].cpu().numpy().astype(np.float64)), logit(r["pred"].astype(np.float64))
    err = float(np.abs(lo - lr_).max())
    assert err <= 2e-2 * float(np.abs(lr_).max()), (err, float(np.abs(lr_).max()))
    assert abs(bce - float(r["bce"])) <= 2e-2 * float(r["bce"])
    assert abs(reg - float(r["reg"])) <= 1e-4 * float(r["reg"])


def test_bf16_auc_tracks_fp32_after_training():
    """north_star: bf16 tensor-core runs agree with fp32 to 1e-3 absolute AUC after a fixed number of steps (150).  The bar is the
    stated 1e-3, checked per seed on a 131 072-row evaluation set (a 16 384-row set adds ~2e-3 of sampling noise to the DIFFERENCE
    of two slightly different rankings, which is what the first round's widened bar had absorbed) and on the mean over seeds."""
    def run(precision, seed):
        class Cfg:
            use_atten = False; use_dcn = False; cdcmdr_precision = precision
        torch.manual_seed(seed)
        rng = np.random.default_rng(seed)
        F, E, T, B = 16, 16, 4, 16384
        fd = np.full(F, 200, dtype=np.int64)
        m = cm.PLE(fd, E, T, 2, 2, ((128, 64), (32,)), (32, 16), dropout=0.0, config=Cfg(), l2_reg_embedding=1e-7,
                   l2_reg_linear=1e-7, l2_reg_dnn=1e-7).to("cuda").train()
        opt = cm.Adam(m.parameters(), lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)
        wtrue = rng.standard_normal((F, 200)) * 1.5

        def batch(n=B):
            x = rng.integers(0, 200, size=(n, F)).astype(np.int32)
            s = wtrue[np.arange(F)[None, :], x].sum(1) / np.sqrt(F) - 1.0
            y = (rng.random(n) < 1 / (1 + np.exp(-2 * s))).astype(np.int16)
            g = rng.integers(0, T, size=n).astype(np.int64)
            return x, y, g
        for _ in range(150):
            x, y, g = batch()
            m.train_step(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), opt, mode="gather", sel=torch.from_numpy(g).cuda())
        m.eval()
        ps, ys = [], []
        for _ in range(8):
            x, y, g = batch()
            with torch.no_grad():
                ps.append(m(torch.from_numpy(x).cuda()).cpu().numpy()[np.arange(B), g])
            ys.append(y)
        return _auc(np.concatenate(ps), np.concatenate(ys))
    diffs = []
    for seed in (5, 6, 7):
        a32, a16 = run("fp32", seed), run("bf16", seed)
        assert a32 > 0.6, a32                               # the model learned something
        diffs.append(a16 - a32)
    assert max(abs(d) for d in diffs) <= 1e-3, diffs
