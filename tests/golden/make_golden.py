#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run HERE (the dev container), where /root/reference exists:

    python tests/golden/make_golden.py

The reference is imported from /root/reference (never copied).  Modules it needs but that
are absent in this image (matplotlib, dataset.aliccp.preprocess_ali_ccp) are stubbed in
sys.modules, exactly as SURVEY.md §8(c) describes.  Every case runs the reference model in
fp32 on CPU with dropout=0 (the reference's dropout RNG cannot be reproduced), seed 2000,
for a few steps of the reference's own training loop body (run.py:483-492 / 635-640):

    pred = model(X...) ; loss = BCELoss(pred_sel, y) + model.get_regularization_loss()
    model.zero_grad(); loss.backward(); Adam(lr=1e-3, betas=(.9,.99), eps=1e-8, wd=1e-8).step()

and dumps inputs, the initial state_dict, per-step predictions / losses / gradients and the
state_dict after every step into one small .npz per case.  The fixtures are what pins
oracle/ (tests/test_oracle_golden.py) — the GPU box has no /root/reference.
"""
import os
import sys
import types
import tempfile

import numpy as np
import torch

REF = os.environ.get("CDCMDR_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def _shim():
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    ds = types.ModuleType("dataset"); al = types.ModuleType("dataset.aliccp")
    pp = types.ModuleType("dataset.aliccp.preprocess_ali_ccp"); pp.reduce_mem = lambda df: df
    sys.modules.setdefault("dataset", ds); sys.modules.setdefault("dataset.aliccp", al)
    sys.modules.setdefault("dataset.aliccp.preprocess_ali_ccp", pp)
    sys.path.insert(0, REF)


_shim()
from model.layer import FeaturesEmbedding, CrossNetV2, CrossNetwork, CrossNetMix  # noqa: E402
from model.ple import PLE  # noqa: E402
from model.mmoe import MMoE  # noqa: E402
from model.dcn import DCN  # noqa: E402
from model.dcnv2 import DCNv2  # noqa: E402
from model.star import STAR  # noqa: E402
from model.cdc import CDC  # noqa: E402


class Cfg:
    use_atten = False
    use_dcn = False
    dataset_name = "synthetic"
    mmoe_n_expert = 3
    ple_n_expert_specific = 2
    ple_n_expert_shared = 1
    p_weight = 0.1
    p_weight_method = "linear_decay"
    p_weight_exp_decay = 0.9
    old_matrix_weight = 0.0
    affinity_func = "minus"


FIELD_DIMS = np.array([7, 5, 11, 4, 9, 6], dtype=np.int64)
E = 4
L2 = dict(l2_reg_embedding=1e-3, l2_reg_linear=1e-3, l2_reg_dnn=1e-3, l2_reg_cross=1e-3)


def make_batch(rng, B, n_group, domain_idx=3):
    x = np.stack([rng.integers(0, d, size=B) for d in FIELD_DIMS], axis=1).astype(np.int32)
    y = (rng.random(B) < 0.35).astype(np.int16).reshape(B, 1)
    g = (x[:, domain_idx] % n_group).astype(np.int64).reshape(B, 1)
    return x, y, g


def sd_np(model):
    return {k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}


def run_case(name, model, fwd, batches, steps, extra=None):
    """fwd(model, x, y, g) -> (pred_selected (B,), target (B,), full_pred tensor)."""
    out = {}
    for k, v in sd_np(model).items():
        out["sd0." + k] = v
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)
    crit = torch.nn.BCELoss()
    model.train()
    for s in range(steps):
        x, y, g = batches[s % len(batches)]
        out[f"in{s}.x"], out[f"in{s}.y"], out[f"in{s}.g"] = x, y, g
        xt, yt, gt = torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(g)
        sel, tgt, full = fwd(model, xt, yt, gt)
        bce = crit(sel, tgt)
        reg = model.get_regularization_loss(device="cpu")
        loss = bce + reg
        model.zero_grad()
        loss.backward()
        out[f"step{s}.pred"] = full.detach().numpy().copy()
        out[f"step{s}.bce"] = np.float32(bce.item())
        out[f"step{s}.reg"] = np.float32(reg.item())
        out[f"step{s}.loss"] = np.float32(loss.item())
        if s == 0:
            for k, p in model.named_parameters():
                if p.grad is not None:
                    out["grad0." + k] = p.grad.detach().numpy().copy()
        opt.step()
        for k, v in sd_np(model).items():
            out[f"sd{s + 1}." + k] = v
    # eval-mode forward on the first batch with the final weights (running stats path)
    model.eval()
    with torch.no_grad():
        x, y, g = batches[0]
        _, _, full = fwd(model, torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(g))
        out["eval.pred"] = full.numpy().copy()
    if extra:
        out.update(extra)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {len(out)} arrays, {os.path.getsize(path) / 1024:.1f} KiB")


def fwd_multi(model, x, y, g):          # run.py:481-484
    p = model(x)
    return p.gather(1, g).squeeze(1), y.squeeze().float(), p


def fwd_single(model, x, y, g):         # run.py:485-488
    p = model(x)
    return p.squeeze(), y.squeeze().float(), p


def fwd_star_grouped(model, x, y, g):   # run.py:477-480
    p, yy = model(x, g, targets=y)
    return p.squeeze(), yy.squeeze().float(), torch.cat([p, yy.float()], dim=1)


def main():
    torch.manual_seed(2000)
    torch.set_num_threads(1)
    rng = np.random.default_rng(2000)
    B, T = 24, 3
    batches = [make_batch(rng, B, T) for _ in range(2)]
    cfg = Cfg()

    # a1: embedding gather (bit-exact)
    emb = FeaturesEmbedding(FIELD_DIMS, E)
    x = torch.from_numpy(batches[0][0])
    np.savez_compressed(os.path.join(OUT, "embedding.npz"), x=batches[0][0], field_dims=FIELD_DIMS,
                        table=emb.embedding_dict.weight.detach().numpy(),
                        out3d=emb(x).detach().numpy(), out2d=emb(x, squeeze_dim=True).detach().numpy())

    # a5/a6: PLE, two CGC levels
    m = PLE(FIELD_DIMS, E, T, 2, 1, ((16, 8), (8,)), (8, 4), dropout=0.0, config=cfg, **L2)
    run_case("ple", m, fwd_multi, batches, 3)

    # a9: MMoE
    m = MMoE(FIELD_DIMS, E, T, 3, (16, 8), (8, 4), dropout=0.0, config=cfg, **L2)
    run_case("mmoe", m, fwd_multi, batches, 3)

    # a10/a13: DCN (rank-1 cross)
    m = DCN(FIELD_DIMS, E, 3, (16, 8), dropout=0.0, **L2)
    run_case("dcn", m, fwd_single, batches, 3)

    # a12/a13: DCNv2 default (CrossNetMix), parallel and stacked
    m = DCNv2(FIELD_DIMS, E, 2, (16, 8), dropout=0.0, low_rank=4, num_experts=3, **L2)
    with torch.no_grad():
        for b in m.crossnet.bias:
            b.normal_(0, 0.1)
    run_case("dcnv2_mix_parallel", m, fwd_single, batches, 3)
    m = DCNv2(FIELD_DIMS, E, 2, (16, 8), dropout=0.0, model_structure="stacked", low_rank=4, num_experts=3, **L2)
    run_case("dcnv2_mix_stacked", m, fwd_single, batches, 2)

    # a10/a11/a12 bare cross layers (CrossNetV2 is unreachable through DCNv2)
    D = len(FIELD_DIMS) * E
    xin = torch.randn(B, D)
    for nm, layer in (("crossv1", CrossNetwork(D, 3)), ("crossv2", CrossNetV2(D, 3)), ("crossmix", CrossNetMix(D, 2, 4, 3))):
        with torch.no_grad():
            for p in (layer.b if hasattr(layer, "b") else layer.bias):
                p.normal_(0, 0.1)
        xi = xin.clone().requires_grad_(True)
        o = layer(xi)
        w = torch.randn_like(o)
        (o * w).sum().backward()
        d = {"x": xin.numpy(), "out": o.detach().numpy(), "w": w.numpy(), "dx": xi.grad.numpy()}
        for k, p in layer.named_parameters():
            d["p." + k] = p.detach().numpy(); d["g." + k] = p.grad.numpy()
        np.savez_compressed(os.path.join(OUT, f"layer_{nm}.npz"), **d)

    # a14: STAR, all-tower mode and grouped (routed) mode
    m = STAR(FIELD_DIMS, E, T, (16, 8), domain_idx=3, dropout=0.0, config=cfg, device="cpu", **L2)
    with torch.no_grad():  # make the shared / domain affine non-trivial
        m.shared_bn_weight.uniform_(0.5, 1.5); m.shared_bn_bias.normal_(0, 0.1)
        for dn in m.domain_norm:
            dn.weight.uniform_(0.5, 1.5); dn.bias.normal_(0, 0.1)
    run_case("star", m, fwd_multi, batches, 3)
    m = STAR(FIELD_DIMS, E, T, (16, 8), domain_idx=3, dropout=0.0, config=cfg, device="cpu", **L2)
    run_case("star_grouped", m, fwd_star_grouped, batches, 3)

    # a15/a16: CDC over PLE (nested expert_dims), MMoE and STAR; three forward modes
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())  # CDC.__init__ makedirs('result/...')
    try:
        n_domain = int(FIELD_DIMS[3])
        d2g = [0, 2, 1, 2]
        w = [0.4, 0.3, 0.2, 0.1]
        for base, ed, td in (("ple", ((16, 8), (8,)), (8, 4)), ("mmoe", (16, 8), (8, 4)), ("star", None, (16, 8))):
            def mk():
                c = CDC(FIELD_DIMS, E, T, n_domain, base, ed, td, 3, domain_cnt_weight=w, n_causal_mask=5,
                        device="cpu", dropout=0.0, config=cfg, **L2)
                c.domain2group_list = list(d2g)
                c.domain2group = torch.tensor(d2g, dtype=torch.int64)
                return c
            dom = 1
            run_case(f"cdc_{base}_warmup", mk(), lambda mm, x, y, g: (lambda p: (p, y.squeeze().float(), p))(mm(x, mode="warmup")),
                     batches, 2, extra={"d2g": np.array(d2g)})
            run_case(f"cdc_{base}_split_domain", mk(),
                     lambda mm, x, y, g: (lambda p: (p, y.squeeze().float(), p))(mm(x, mode="split", domain_i=dom)),
                     batches, 2, extra={"d2g": np.array(d2g), "domain_i": np.array(dom)})
            run_case(f"cdc_{base}_split_gather", mk(),
                     lambda mm, x, y, g: (lambda p: (p.squeeze(1), y.squeeze().float(), p))(mm(x, mode="split")),
                     batches, 2, extra={"d2g": np.array(d2g)})
        # N4 known-answer: causal kernel
        X = np.array([[.1, .2, .3], [.4, .1, .2], [.3, .3, .9], [.5, .7, .2]])
        kap = CDC.calc_causal_matrix(X)
        rr = np.random.default_rng(7).random((6, 9))
        np.savez_compressed(os.path.join(OUT, "causal_matrix.npz"), X=X, kappa=np.asarray(kap), X2=rr,
                            kappa2=np.asarray(CDC.calc_causal_matrix(rr)))
    finally:
        os.chdir(cwd)


if __name__ == "__main__":
    main()
