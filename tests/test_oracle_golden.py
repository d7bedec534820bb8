"""Pin oracle/ against fixtures produced by the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import cdcmdr_oracle as O
from tests.golden_cases import ALL_CASES as CASES, ATTEN, FIELD_DIMS, E, L2, load, state, strip

KINDS = {"ple": O.PLE, "mmoe": O.MMoE, "dcn": O.DCN, "dcnv2": O.DCNv2, "star": O.STAR, "autoint": O.AutoInt}
RTOL, ATOL = 1e-4, 2e-6      # north_star: fp32 logits and gradients within 1e-4 relative


def bias_before_bn(kind, k):
    if kind == "star" and k == "shared_bn_bias":
        return True
    if not k.endswith(".bias"):
        return False
    if kind in ("ple", "mmoe") and k.startswith(("towers.", "experts.")):
        return int(k.split(".")[-2]) % 4 == 0 and not k.endswith("layers.8.bias")
    if kind in ("dcn", "dcnv2", "autoint") and k.startswith(("mlp.", "dnn.")):
        return int(k.split(".")[-2]) % 4 == 0
    if kind == "star":      # PN beta (shared and per-domain) is also cancelled by the BN after the first FC
        return ".linears." in k or k == "shared_bn_bias" or k.startswith("domain_norm.")
    return False


def close(a, b, what, rtol=RTOL, atol=ATOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    assert (err <= tol).all(), f"{what}: max err {err.max():.3e} at tol {tol.flat[err.argmax()]:.3e}"


def test_embedding_bit_exact():
    g = load("embedding")
    off = O.field_offsets(g["field_dims"])
    out, idx = O.embed_gather(g["x"], off, g["table"])
    assert np.array_equal(out, g["out2d"])
    assert np.array_equal(out.reshape(g["out3d"].shape), g["out3d"])


@pytest.mark.parametrize("name", sorted(CASES))
def test_train_steps_match_reference(name):
    kind, kw, mode, steps = CASES[name]
    gold = load(name)
    model = KINDS[kind](FIELD_DIMS, E, **kw, **{k: v for k, v in L2.items() if kind != "autoint" or k != "l2_reg_cross"})
    if name in ATTEN:                                        # config.use_atten (tests/golden/make_golden_atten.py)
        model.enable_atten(**ATTEN[name])
    sd = state(gold, 0)
    opt = O.Adam()
    sel = {}
    if "d2g" in gold:
        sel = dict(domain2group=gold["d2g"], domain_idx=3)
        if "domain_i" in gold:
            sel["domain_i"] = int(gold["domain_i"])
    for s in range(steps):
        x, y, g = gold[f"in{s}.x"], gold[f"in{s}.y"], gold[f"in{s}.g"]
        if mode in ("gather", "star_grouped"):
            sel["group"] = g
        r = O.train_step(model, sd, opt, x, y, mode, **sel)
        if mode == "star_grouped":
            close(r["pred"], gold[f"step{s}.pred"][:, :1], f"{name} step{s} pred")
            perm, _ = O.route_partition(g, kw["n_tower"])
            assert np.array_equal(y.reshape(-1)[perm], gold[f"step{s}.pred"][:, 1].astype(np.int64)), "routing order"
        elif name.startswith("cdc_"):
            close(r["psel"], gold[f"step{s}.pred"].reshape(-1), f"{name} step{s} pred")
        else:
            close(r["pred"], gold[f"step{s}.pred"].reshape(r["pred"].shape), f"{name} step{s} pred")
        close(r["bce"], gold[f"step{s}.bce"], "bce"); close(r["reg"], gold[f"step{s}.reg"], "reg")
        if s == 0:
            gk = {strip(k[6:]): v for k, v in gold.items() if k.startswith("grad0.")}
            assert set(gk) == set(r["grads"]), set(gk) ^ set(r["grads"])
            for k, v in gk.items():
                close(r["grads"][k].reshape(v.shape), v, f"{name} grad {k}", atol=2e-6 if bias_before_bn(kind, k) else 1e-7)
        ref = state(gold, s + 1)
        assert set(ref) == set(sd)
        for k, v in ref.items():
            # a Linear bias that feeds BatchNorm has an exactly-zero true gradient; what Adam sees is
            # fp32 rounding noise normalised to +-lr per step, in the reference as much as here.
            if k.endswith("in_proj_bias"):
                # the key bias of an attention layer has a mathematically zero gradient (softmax cancels a shift common to a row's
                # scores): rounding noise -> +-lr steps under Adam, like a bias in front of a batch-statistics BatchNorm
                n3 = v.shape[0] // 3
                close(sd[k][n3:2 * n3], v[n3:2 * n3], f"{name} step{s + 1} {k} (key third)", atol=2.1e-3 * (s + 1))
                close(np.delete(sd[k], np.s_[n3:2 * n3]), np.delete(v, np.s_[n3:2 * n3]), f"{name} step{s + 1} {k}", atol=1e-6)
                continue
            close(sd[k].reshape(v.shape), v, f"{name} step{s + 1} {k}",
                  atol=2.1e-3 * (s + 1) if bias_before_bn(kind, k) else (3e-4 * (s + 1) if k.endswith("running_mean") else 1e-6))
    # eval-mode forward with the reference's final weights (eval BN does not cancel the noise-driven biases)
    sd = state(gold, steps)
    fkw = {"group": gold["in0.g"]} if mode == "star_grouped" else {}
    pred, _ = model.forward(sd, gold["in0.x"], train=False, **fkw)
    if mode == "star_grouped":
        close(pred, gold["eval.pred"][:, :1], "eval pred")
    elif mode in ("gather", "single"):
        close(pred, gold["eval.pred"].reshape(pred.shape), "eval pred")
    else:
        psel, _ = O.select_pred(pred, mode, x=gold["in0.x"], **sel)
        close(psel, gold["eval.pred"].reshape(psel.shape), "eval pred")


@pytest.mark.parametrize("nm", ["crossv1", "crossv2", "crossmix"])
def test_bare_cross_layers(nm):
    g = load("layer_" + nm)
    sd = {"c." + k[2:]: v for k, v in g.items() if k.startswith("p.")}
    grads = {}
    if nm == "crossv1":
        out, c = O.cross_v1_fwd(sd, "c", 3, g["x"]); dx = O.cross_v1_bwd(sd, "c", 3, c, g["w"], grads)
    elif nm == "crossv2":
        out, c = O.cross_v2_fwd(sd, "c", 3, g["x"]); dx = O.cross_v2_bwd(sd, "c", 3, c, g["w"], grads)
    else:
        out, c = O.cross_mix_fwd(sd, "c", 2, 3, g["x"]); dx = O.cross_mix_bwd(sd, "c", 2, 3, g["x"], c, g["w"], grads)
    close(out, g["out"], "out"); close(dx, g["dx"], "dx", atol=1e-5)
    for k, v in g.items():
        if k.startswith("g."):
            close(grads["c." + k[2:]].reshape(v.shape), v, k, atol=1e-5)


def test_causal_matrix_known_answer():
    g = load("causal_matrix")
    close(O.calc_causal_matrix(g["X"]), g["kappa"], "kappa", rtol=1e-9, atol=1e-12)
    close(O.calc_causal_matrix(g["X2"]), g["kappa2"], "kappa2", rtol=1e-9, atol=1e-12)
    # SURVEY §4 probe values
    k = O.calc_causal_matrix(g["X"])
    assert abs(k[0, 1] - 0.092641) < 1e-6 and abs(k[1, 2] - 0.714029) < 1e-6


def test_routing_order_known_answer():
    # SURVEY §3.5 probe
    perm, counts = O.route_partition([0, 1, 0, 3, 1, 0, 2, 1, 3, 3, 0, 1], 4)
    assert perm.tolist() == [0, 2, 5, 10, 1, 4, 7, 11, 6, 3, 8, 9] and counts.tolist() == [4, 4, 1, 3]
