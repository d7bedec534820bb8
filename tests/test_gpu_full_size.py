"""BASELINE.json configs at their full shapes on the B200.  Against the CPU oracle itself (one full-size oracle step takes 5-15 s
of numpy): C1 (fp32), C2 DCNv2 with CrossNetMix and CrossNetV2, C3 PLE / MMoE and C4 CDC-PLE - fp32 path at 1e-4, bf16 tensor-core
path at 2e-2 on the logit scale (north_star) - `test_*_matches_oracle`.  And through size-independent properties: bit-repeatability
of the whole fused step (every reduction has a fixed order), CUDA-graph replay == eager, row routing == all-rows computation +
selection (bit-exact order)."""
import numpy as np
import pytest
import torch

import cdcmdr_b200 as cm
from oracle import cdcmdr_oracle as O

pytestmark = pytest.mark.gpu
L2 = dict(l2_reg_embedding=1e-5, l2_reg_linear=1e-5, l2_reg_dnn=1e-5)
ADAM = dict(lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)


def cfg(precision="fp32", **kw):
    class Cfg:
        use_atten = False; use_dcn = False; ple_n_expert_specific = 2; ple_n_expert_shared = 2; mmoe_n_expert = 8
        cdcmdr_precision = precision
    for k, v in kw.items():
        setattr(Cfg, k, v)
    return Cfg()


def data(fd, B, seed, zipf=True):
    rng = np.random.default_rng(seed)
    cols = [np.minimum(rng.zipf(1.05, size=B) - 1, d - 1) if zipf else rng.integers(0, d, size=B) for d in fd]
    x = np.stack(cols, axis=1).astype(np.int32)
    y = (rng.random(B) < 0.05).astype(np.int16)
    return x, y, rng


def logit(p):
    p = np.clip(p.astype(np.float64), 1e-12, 1 - 1e-12)
    return np.log(p / (1 - p))


def snapshot(model):
    return {k: v.detach().clone() for k, v in model.state_dict().items()}


def test_c1_ple_fp32_matches_oracle():
    """configs[0]: PLE fp32, 4 domains, 16 fields x embed 16, vocab 1M, batch 2048 - the reference's own CPU-runnable case."""
    F, E, T, B = 16, 16, 4, 2048
    fd = np.full(F, 66_666, dtype=np.int64); fd[-1] = 4
    dims, tower = ((256, 128), (64,)), (64, 32)
    torch.manual_seed(2000)
    model = cm.PLE(fd, E, T, 2, 2, dims, tower, dropout=0.0, config=cfg(), **L2)
    sd = {k: (v.detach().numpy().astype(np.float64) if v.dtype == torch.float32 else v.detach().numpy().copy())
          for k, v in model.state_dict().items()}
    om = O.PLE(fd, E, T, 2, 2, dims, tower, **L2)
    x, y, rng = data(fd, B, 1)
    g = x[:, -1].astype(np.int64)
    model = model.to("cuda").train()
    opt, oopt = cm.Adam(model.parameters(), **ADAM), O.Adam()
    xt, yt, gt = (torch.from_numpy(a).cuda() for a in (x, y, g))
    for s in range(2):
        r = O.train_step(om, sd, oopt, x, y, "gather", group=g)
        out = model.train_step(xt, yt, opt, mode="gather", sel=gt)
        loss, bce, reg = model.step_losses(out)
        assert np.abs(out["pred"].cpu().numpy() - r["pred"]).max() <= (1e-4 if s == 0 else 3e-3)
        assert abs(bce - float(r["bce"])) <= 1e-4 * float(r["bce"]) + (0 if s == 0 else 1e-3)
        assert abs(reg - float(r["reg"])) <= 1e-4 * float(r["reg"])
    tab = model.state_dict()["embedding.embedding_dict.weight"].cpu().numpy()
    d = np.abs(tab - sd["embedding.embedding_dict.weight"])
    assert d.max() <= 2.1e-3 * 2 and (d > 1e-5).mean() < 1e-3        # all 1M rows took the dense Adam step


def _same_state(a, b):
    return all(torch.equal(a[k], b[k]) for k in a)


def test_c2_dcnv2_bf16_full_shape():
    """configs[1]: DCNv2, 3 cross layers + MLP 512-256-128, 26 fields x embed 32, batch 16384, bf16."""
    F, E, B = 26, 32, 16384
    fd = np.full(F, 40_000, dtype=np.int64)
    x, y, rng = data(fd, B, 2)
    xt, yt = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()

    def build(precision, mix):
        torch.manual_seed(7)
        return cm.DCNv2(fd, E, 3, (512, 256, 128), dropout=0.0, use_low_rank_mixture=mix, config=cfg(precision), **L2).to("cuda")
    for mix in (True, False):                                        # CrossNetMix (stock) and CrossNetV2 (north_star formula)
        m32, m16 = build("fp32", mix).eval(), build("bf16", mix).eval()
        with torch.no_grad():
            p32, p16 = m32(xt).cpu().numpy(), m16(xt).cpu().numpy()
        assert p32.shape == (B,)
        l32 = logit(p32)
        assert np.abs(logit(p16) - l32).max() <= 2e-2 * max(1.0, np.abs(l32).max())
        runs = []
        for _ in range(2):                                           # the fused training step is bit-repeatable
            m = build("bf16", mix).train()
            opt = cm.Adam(m.parameters(), **ADAM)
            losses = [m.step_losses(m.train_step(xt, yt, opt, mode="col", col=0))[1] for _ in range(3)]
            runs.append((losses, snapshot(m)))
        assert runs[0][0] == runs[1][0] and _same_state(runs[0][1], runs[1][1])
        assert np.isfinite(runs[0][0]).all() and runs[0][0][-1] < runs[0][0][0]


@pytest.mark.parametrize("kind", ["ple", "mmoe"])
def test_c3_eight_experts_bf16_full_batch(kind):
    """configs[2]: 8 experts (PLE: 3 tasks x 2 specific + 2 shared, 2 levels; MMoE: 8), 3 tasks, 10 domains, batch 65536, bf16."""
    F, E, T, B = 23, 16, 3, 65536
    fd = np.full(F, 45_000, dtype=np.int64); fd[10] = 10
    x, y, rng = data(fd, B, 3)
    g = (x[:, 10] % T).astype(np.int64)
    xt, yt, gt = (torch.from_numpy(a).cuda() for a in (x, y, g))

    def build(precision):
        torch.manual_seed(9)
        if kind == "ple":
            return cm.PLE(fd, E, T, 2, 2, ((256, 128), (64,)), (64, 32), dropout=0.0, config=cfg(precision), **L2).to("cuda")
        return cm.MMoE(fd, E, T, 8, (256, 128, 64), (64, 32), dropout=0.0, config=cfg(precision), **L2).to("cuda")
    m32, m16 = build("fp32").eval(), build("bf16").eval()
    with torch.no_grad():
        l32, l16 = logit(m32(xt).cpu().numpy()), logit(m16(xt).cpu().numpy())
    assert np.abs(l16 - l32).max() <= 2e-2 * max(1.0, np.abs(l32).max())
    runs = []
    for _ in range(2):
        m = build("bf16").train()
        opt = cm.Adam(m.parameters(), **ADAM)
        losses = [m.step_losses(m.train_step(xt, yt, opt, mode="gather", sel=gt))[1] for _ in range(3)]
        runs.append((losses, snapshot(m)))
    assert runs[0][0] == runs[1][0] and _same_state(runs[0][1], runs[1][1])
    assert runs[0][0][-1] < runs[0][0][0]


def test_c4_cdc_ple_graph_replay_equals_eager():
    """configs[3]: CDC over 30 domains on a PLE backbone, batch 65536: the captured CUDA graph of the fused step and the eager
    step produce bit-identical parameters and losses (dropout on: the mask is a function of the device-resident step state)."""
    F, E, T, nd, B = 23, 16, 4, 30, 65536
    fd = np.full(F, 45_000, dtype=np.int64); fd[10] = nd
    x, y, rng = data(fd, B, 4)
    x[:, 10] = 7
    xt, yt = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()

    def build():
        torch.manual_seed(2000)
        m = cm.CDC(fd, E, T, nd, "ple", ((256, 128), (64,)), (64, 32), 10, dropout=0.2, config=cfg("bf16"), **L2).to("cuda").train()
        m.set_groups([d % T for d in range(nd)])
        return m, cm.Adam(m.parameters(), **ADAM)
    m1, o1 = build()
    eager = [m1.step_losses(m1.train_step(xt, yt, o1, mode="split", domain_i=7)) for _ in range(5)]
    m2, o2 = build()
    gs = cm.GraphedTrainStep(m2, o2, B, F, mode="split", domain_i=7, warmup=2)
    gs.x.copy_(xt); gs.y.copy_(yt)
    gs.capture()                                                     # 2 real warm-up steps ran before the capture
    graphed = [m2.step_losses(gs()) for _ in range(3)]
    assert graphed == eager[2:5]
    assert _same_state(snapshot(m1), snapshot(m2))
    assert gs.launches_per_step > 50


def test_c5_star_routing_equals_selection():
    """configs[4] shape (embed 64, towers 256-128-64-32, batch 65536; vocabulary scaled to fit a test): with running statistics
    (eval) a tower's output for a row does not depend on which other rows it sees, so the row-routed forward must equal the
    all-rows forward followed by per-row tower selection, in the partition's bit-exact order."""
    F, E, T, nd, B = 23, 64, 4, 30, 65536
    fd = np.full(F, 20_000, dtype=np.int64); fd[10] = nd
    x, y, rng = data(fd, B, 5)
    x[:, 10] = rng.integers(0, nd, size=B)
    group = (x[:, 10] % (T + 1) - 1).astype(np.int64)                 # -1: rows outside every group are dropped
    torch.manual_seed(3)
    m = cm.STAR(fd, E, T, (256, 128, 64, 32), domain_idx=10, dropout=0.0, config=cfg("fp32"), **L2).to("cuda")
    with torch.no_grad():                                            # non-trivial normalisation statistics
        for t in range(T):
            m.domain_norm[t].running_mean.normal_(0, 0.3); m.domain_norm[t].running_var.uniform_(0.5, 1.5)
    m.eval()
    xt, gt, yt = torch.from_numpy(x).cuda(), torch.from_numpy(group).cuda(), torch.from_numpy(y).cuda()
    with torch.no_grad():
        full = m(xt).cpu().numpy()
        routed, yperm = m(xt, gt.view(-1, 1), targets=yt.view(-1, 1))
    perm = np.concatenate([np.flatnonzero(group == t) for t in range(T)])
    assert routed.shape == (len(perm), 1)
    assert np.array_equal(yperm.cpu().numpy().reshape(-1), y[perm])
    np.testing.assert_allclose(routed.cpu().numpy()[:, 0], full[perm, group[perm]], rtol=0, atol=2e-6)
    # and the CDC wrapper over STAR selects towers per sample bit-exactly
    torch.manual_seed(3)
    c = cm.CDC(fd, E, T, nd, "star", None, (256, 128, 64, 32), 10, dropout=0.0, config=cfg("fp32"), **L2).to("cuda").eval()
    c.set_groups([d % T for d in range(nd)])
    with torch.no_grad():
        cf = c.base_model_instance(xt).cpu().numpy()
        cs = c(xt, mode="split").cpu().numpy()[:, 0]
    assert np.array_equal(cs, cf[np.arange(B), np.array(c.domain2group_list)[x[:, 10]]])


# ------------------------------------------------------------------------------------------------ full shapes against the oracle
def _sd64(model):
    return {k: (v.detach().cpu().numpy().astype(np.float64) if v.dtype == torch.float32 else v.detach().cpu().numpy().copy())
            for k, v in model.state_dict().items()}


def _check_step_vs_oracle(build, om, x, y, step_kw, oracle_kw, strip=""):
    """One fused training step of the fp32 path and of the bf16 path against ONE oracle step (float64 numpy, reference
    arithmetic) on identical weights and inputs: predictions, BCE, regulariser."""
    m32 = build("fp32")
    sd = {k[len(strip):]: v for k, v in _sd64(m32).items()}
    r = O.train_step(om, sd, O.Adam(), x, y, **oracle_kw)
    ref_p = np.asarray(r["pred"], dtype=np.float64)
    lr_ = logit(ref_p)
    xt, yt = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    for prec, model in (("fp32", m32), ("bf16", None)):
        if model is None:
            model = build(prec)
        model = model.to("cuda").train()
        opt = cm.Adam(model.parameters(), **ADAM)
        out = model.train_step(xt, yt, opt, **step_kw(model))
        loss, bce, reg = model.step_losses(out)
        p = out["pred"].cpu().numpy().astype(np.float64).reshape(ref_p.shape)
        if prec == "fp32":
            assert np.abs(p - ref_p).max() <= 1e-4, (prec, float(np.abs(p - ref_p).max()))
            assert abs(bce - float(r["bce"])) <= 1e-4 * float(r["bce"]), (prec, bce, float(r["bce"]))
        else:
            err = float(np.abs(logit(p) - lr_).max())
            assert err <= 2e-2 * max(1.0, float(np.abs(lr_).max())), (prec, err, float(np.abs(lr_).max()))
            assert abs(bce - float(r["bce"])) <= 2e-2 * float(r["bce"]), (prec, bce, float(r["bce"]))
        assert abs(reg - float(r["reg"])) <= 1e-4 * float(r["reg"]), (prec, reg, float(r["reg"]))
        del model, opt, out
        torch.cuda.empty_cache()


def test_c4_cdc_ple_full_batch_matches_oracle():
    """configs[3] at its full shape (30 domains -> 4 clusters, 23 fields x 16, vocab 1M, batch 65536), mode='split' with the
    per-sample tower gather, dropout 0: fp32 path and bf16 path against the oracle."""
    F, E, T, nd, B = 23, 16, 4, 30, 65536
    fd = np.full(F, 45_000, dtype=np.int64); fd[10] = nd
    x, y, rng = data(fd, B, 14)
    x[:, 10] = rng.integers(0, nd, size=B)
    d2g = [d % T for d in range(nd)]

    def build(precision):
        torch.manual_seed(2000)
        m = cm.CDC(fd, E, T, nd, "ple", ((256, 128), (64,)), (64, 32), 10, dropout=0.0, config=cfg(precision), **L2)
        m.set_groups(d2g)
        return m
    om = O.PLE(fd, E, T, 2, 2, ((256, 128), (64,)), (64, 32), **L2)
    _check_step_vs_oracle(build, om, x, y, lambda m: dict(mode="split", domain_i=None),
                          dict(mode="split_gather", domain2group=np.array(d2g), domain_idx=10), strip="base_model_instance.")


@pytest.mark.parametrize("kind", ["ple", "mmoe"])
def test_c3_eight_experts_full_batch_matches_oracle(kind):
    """configs[2] at its full shape (8 experts, 3 tasks, batch 65536): fp32 and bf16 paths against the oracle."""
    F, E, T, B = 23, 16, 3, 65536
    fd = np.full(F, 45_000, dtype=np.int64); fd[10] = 10
    x, y, rng = data(fd, B, 13)
    g = (x[:, 10] % T).astype(np.int64)
    gt = torch.from_numpy(g).cuda()

    def build(precision):
        torch.manual_seed(9)
        if kind == "ple":
            return cm.PLE(fd, E, T, 2, 2, ((256, 128), (64,)), (64, 32), dropout=0.0, config=cfg(precision), **L2)
        return cm.MMoE(fd, E, T, 8, (256, 128, 64), (64, 32), dropout=0.0, config=cfg(precision), **L2)
    om = (O.PLE(fd, E, T, 2, 2, ((256, 128), (64,)), (64, 32), **L2) if kind == "ple"
          else O.MMoE(fd, E, T, 8, (256, 128, 64), (64, 32), **L2))
    _check_step_vs_oracle(build, om, x, y, lambda m: dict(mode="gather", sel=gt), dict(mode="gather", group=g))


@pytest.mark.parametrize("mix", [True, False])
def test_c2_dcnv2_full_shape_matches_oracle(mix):
    """configs[1] at its full shape (26 fields x 32, 3 cross layers, MLP 512-256-128, batch 16384): CrossNetMix (stock) and
    CrossNetV2 (the north_star formula), fp32 and bf16 paths against the oracle."""
    F, E, B = 26, 32, 16384
    fd = np.full(F, 40_000, dtype=np.int64)
    x, y, rng = data(fd, B, 12)

    def build(precision):
        torch.manual_seed(7)
        return cm.DCNv2(fd, E, 3, (512, 256, 128), dropout=0.0, use_low_rank_mixture=mix, config=cfg(precision), **L2)
    om = O.DCNv2(fd, E, 3, (512, 256, 128), use_low_rank_mixture=mix, **L2)
    _check_step_vs_oracle(build, om, x, y, lambda m: dict(mode="col", col=0), dict(mode="single"))


def test_c4_stock_attention_block_bf16_matches_fp32_path():
    """configs[3] with the reference's STOCK attention settings (config.py:24-28: 64-d, 2 heads, 3 layers, V_res) at batch 16384:
    the bf16 path (token matrices bf16, projections on the tcgen05 GEMM, bf16 attention core) against the fp32 path on identical
    weights - eval logits 2e-2, first training step's BCE 2e-2 - and the block's parameters train."""
    F, E, T, nd, B = 23, 16, 4, 30, 16384
    fd = np.full(F, 45_000, dtype=np.int64); fd[10] = nd
    x, y, rng = data(fd, B, 15)
    x[:, 10] = rng.integers(0, nd, size=B)
    xt, yt = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    att = dict(use_atten=True, atten_embed_dim=64, att_layer_num=3, att_head_num=2, att_res=True)
    res, sd = {}, None
    for prec in ("fp32", "bf16"):
        torch.manual_seed(2000)
        m = cm.CDC(fd, E, T, nd, "ple", ((256, 128), (64,)), (64, 32), 10, dropout=0.0, config=cfg(prec, **att), **L2)
        if sd is None:
            sd = {k: v.clone() for k, v in m.state_dict().items()}
        else:
            m.load_state_dict(sd, strict=True)
        m = m.to("cuda")
        m.set_groups([d % T for d in range(nd)])
        m.eval()
        with torch.no_grad():
            p = m(xt, mode="split").cpu().numpy()
        m.train()
        opt = cm.Adam(m.parameters(), **ADAM)
        w0 = m.base_model_instance.state_dict()["self_attns.1.in_proj_weight"].clone()
        out = m.train_step(xt, yt, opt, mode="split", domain_i=None)
        bce = m.step_losses(out)[1]
        moved = float((m.base_model_instance.state_dict()["self_attns.1.in_proj_weight"] - w0).abs().max())
        res[prec] = (logit(p), bce, moved)
        del m, opt, out
        torch.cuda.empty_cache()
    l32, l16 = res["fp32"][0], res["bf16"][0]
    assert np.abs(l16 - l32).max() <= 2e-2 * max(1.0, float(np.abs(l32).max())), float(np.abs(l16 - l32).max())
    assert abs(res["bf16"][1] - res["fp32"][1]) <= 2e-2 * abs(res["fp32"][1])
    assert res["fp32"][2] > 1e-4 and res["bf16"][2] > 1e-4
