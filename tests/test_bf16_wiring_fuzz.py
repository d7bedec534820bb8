"""The bf16 tensor-core path's HOST wiring (operand views with padded pitches, the ones column of the gathered embeddings, fused gate
tails, grouped launches, bf16 / fp32 hand-overs around BatchNorm and the attention block) on random model geometries, through the
host-memory emulator of the C-ABI (which rounds to bf16 where the kernels do): the same weights on the fp32 path and on the bf16
path must agree to bf16 accuracy - eval logits within 4e-2 of the logit scale, the first and second fused steps' BCE within
3-4 %.  The GPU tests pin the kernels at a few shapes; this pins the launch wiring across shapes (200 seeds swept offline)."""
import numpy as np
import pytest
import torch

import cdcmdr_b200 as cm
from oracle.host_abi import HostABI

L2 = dict(l2_reg_embedding=1e-3, l2_reg_linear=1e-3, l2_reg_dnn=1e-3, l2_reg_cross=1e-3)
KINDS = ["ple", "mmoe", "star", "dcn", "dcnv2"]


def _config(rng, prec, atten):
    class Cfg:
        pass
    c = Cfg()
    c.use_atten, c.use_dcn, c.cdcmdr_precision = atten, False, prec
    c.mmoe_n_expert, c.ple_n_expert_specific, c.ple_n_expert_shared = int(rng.integers(1, 5)), int(rng.integers(1, 4)), int(rng.integers(1, 4))
    if atten:
        c.att_head_num = int(rng.choice([1, 2]))
        c.atten_embed_dim = c.att_head_num * int(rng.choice([4, 8]))
        c.att_layer_num, c.att_res = int(rng.integers(1, 3)), bool(rng.integers(0, 2))
    return c


def _build(kind, fd, E, T, cfg, seed):
    r = np.random.default_rng(seed)
    hid = lambda: int(r.choice([8, 16, 24, 32]))   # noqa: E731  (the bf16 path needs widths that are multiples of 8)
    if kind == "ple":
        dims = tuple(tuple(hid() for _ in range(int(r.integers(1, 3)))) for _ in range(int(r.integers(1, 4))))
        return cm.PLE(fd, E, T, cfg.ple_n_expert_specific, cfg.ple_n_expert_shared, dims, tuple(hid() for _ in range(int(r.integers(1, 3)))),
                      dropout=0.0, config=cfg, **L2), "multi"
    if kind == "mmoe":
        return cm.MMoE(fd, E, T, cfg.mmoe_n_expert, tuple(hid() for _ in range(int(r.integers(1, 4)))),
                       tuple(hid() for _ in range(int(r.integers(1, 3)))), dropout=0.0, config=cfg, **L2), "multi"
    if kind == "star":
        return cm.STAR(fd, E, T, tuple(hid() for _ in range(int(r.integers(1, 4)))), domain_idx=1, dropout=0.0, config=cfg, device="cpu",
                       **L2), "multi"
    if kind == "dcn":
        return cm.DCN(fd, E, int(r.integers(1, 4)), tuple(hid() for _ in range(int(r.integers(1, 3)))), dropout=0.0, config=cfg, **L2), "single"
    return cm.DCNv2(fd, E, int(r.integers(1, 3)), tuple(hid() for _ in range(int(r.integers(1, 3)))), dropout=0.0,
                    model_structure=str(r.choice(["parallel", "stacked"])), use_low_rank_mixture=bool(r.integers(0, 2)), low_rank=8,
                    num_experts=int(r.integers(1, 4)), config=cfg, **L2), "single"


@pytest.fixture
def emulator():
    old = cm._lib._LIB
    cm._lib.install(HostABI())
    yield
    cm._lib.install(old)


@pytest.mark.parametrize("block", range(2))
def test_bf16_path_wiring_agrees_with_fp32_path(block, emulator):
    for seed in range(20 * block, 20 * block + 20):
        rng = np.random.default_rng(seed)
        kind = KINDS[seed % 5]
        E, F, T = int(rng.choice([8, 16])), int(rng.integers(2, 6)), int(rng.integers(1, 5))
        fd = rng.integers(3, 12, size=F).astype(np.int64)
        fd[1] = max(T, 3)
        B = int(rng.choice([1, 17, 40, 130]))
        atten = kind in ("ple", "mmoe", "star") and bool(rng.integers(0, 2))
        x = torch.from_numpy(np.stack([rng.integers(0, d, size=B) for d in fd], axis=1).astype(np.int32))
        y = torch.from_numpy((rng.random((B, 1)) < 0.4).astype(np.int16))
        g = torch.from_numpy((x[:, 1].numpy() % T).astype(np.int64).reshape(B, 1))
        what = dict(seed=seed, kind=kind, E=E, F=F, T=T, B=B, atten=atten)
        res, sd = {}, None
        for prec in ("fp32", "bf16"):
            cfg = _config(np.random.default_rng(seed + 1000), prec, atten)
            torch.manual_seed(seed)
            m, mode = _build(kind, fd, E, T, cfg, seed)
            if sd is None:
                sd = {k: v.clone() for k, v in m.state_dict().items()}
            else:
                m.load_state_dict(sd, strict=True)
            m.eval()
            with torch.no_grad():
                pe = m(x).numpy().reshape(B, -1).copy()
            m.train()
            opt = cm.Adam(m.parameters(), lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)
            step = (lambda: m.train_step(x, y, opt, mode="gather", sel=g)) if mode == "multi" else (lambda: m.train_step(x, y, opt, mode="col", col=0))
            b1 = m.step_losses(step())[1]
            b2 = m.step_losses(step())[1]
            res[prec] = (pe, b1, b2)
        logit = lambda p: np.log(np.clip(p, 1e-7, 1 - 1e-7)) - np.log1p(-np.clip(p, 1e-7, 1 - 1e-7))   # noqa: E731
        l32, l16 = logit(res["fp32"][0].astype(np.float64)), logit(res["bf16"][0].astype(np.float64))
        assert np.abs(l32 - l16).max() <= 4e-2 * max(1.0, float(np.abs(l32).max())), (what, "eval logits")
        assert abs(res["fp32"][1] - res["bf16"][1]) <= 3e-2 * max(abs(res["fp32"][1]), 1e-3), (what, "first step")
        assert abs(res["fp32"][2] - res["bf16"][2]) <= 4e-2 * max(abs(res["fp32"][2]), 1e-3), (what, "second step")


@pytest.mark.parametrize("dims,ns,nsh", [(((128, 64), (16,)), 2, 1), (((256, 128), (64,)), 1, 2), (((128, 128),), 2, 2)])
def test_ple_chain_kernel_wiring_matches_per_layer_path(dims, ns, nsh, emulator, monkeypatch):
    """PLE level 0 through cdcmdr_ple_chain_fwd (expert layers 0 -> 1 chained in one launch + the gate logits; geometry d0 in
    {128, 256}, d1 in {64, 128}) against the per-layer GEMM launches of the same bf16 path on identical weights, through the
    emulator: eval predictions (where the chain skips the layer-0 activation store), two fused training steps, and the fp32 path
    as the outer reference."""
    rng = np.random.default_rng(7)
    E, F, T, B = 8, 5, 3, 70
    fd = rng.integers(3, 12, size=F).astype(np.int64)
    x = torch.from_numpy(np.stack([rng.integers(0, d, size=B) for d in fd], axis=1).astype(np.int32))
    y = torch.from_numpy((rng.random((B, 1)) < 0.4).astype(np.int16))
    g = torch.from_numpy(rng.integers(0, T, size=(B, 1)).astype(np.int64))
    res, sd = {}, None
    for tag, prec, chain in (("fp32", "fp32", "1"), ("chain", "bf16", "1"), ("layers", "bf16", "0")):
        monkeypatch.setenv("CDCMDR_PLE_CHAIN", chain)

        class Cfg:
            use_atten = False; use_dcn = False; cdcmdr_precision = prec
        torch.manual_seed(3)
        m = cm.PLE(fd, E, T, ns, nsh, dims, (16, 8), dropout=0.0, config=Cfg(), **L2)
        if sd is None:
            sd = {k: v.clone() for k, v in m.state_dict().items()}
        else:
            m.load_state_dict(sd, strict=True)
        assert m._levels[0].chain == (tag == "chain")
        m.eval()
        with torch.no_grad():
            pe = m(x).numpy().copy()
        m.train()
        opt = cm.Adam(m.parameters(), lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)
        b1 = m.step_losses(m.train_step(x, y, opt, mode="gather", sel=g))[1]
        b2 = m.step_losses(m.train_step(x, y, opt, mode="gather", sel=g))[1]
        w = m.state_dict()["cgc_layers.0.experts_specific.0.layers.3.weight"].numpy().copy()
        res[tag] = (pe, b1, b2, w)
    logit = lambda p: np.log(np.clip(p, 1e-7, 1 - 1e-7)) - np.log1p(-np.clip(p, 1e-7, 1 - 1e-7))   # noqa: E731
    lc, ll, l32 = (logit(res[k][0].astype(np.float64)) for k in ("chain", "layers", "fp32"))
    scale = max(1.0, float(np.abs(l32).max()))
    assert np.abs(lc - ll).max() <= 5e-3 * scale                 # same bf16 roundings, different accumulation order
    assert np.abs(lc - l32).max() <= 4e-2 * scale
    for i in (1, 2):
        assert abs(res["chain"][i] - res["layers"][i]) <= 5e-3 * abs(res["layers"][i])
        assert abs(res["chain"][i] - res["fp32"][i]) <= 4e-2 * abs(res["fp32"][i])
    assert np.abs(res["chain"][3] - res["layers"][3]).max() <= 2.1e-3    # two Adam steps: +-lr sign noise at most
