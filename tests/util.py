"""Shared driver for the golden-fixture cases: builds OUR model with the fixture's constructor arguments, loads the
reference's initial state_dict, runs the reference's training-loop body (run.py:483-492 / 635-640) through either the
fused `train_step` or the drop-in autograd path, and compares against what the unmodified reference produced."""
import numpy as np
import torch

import cdcmdr_b200 as cm
from tests.golden_cases import ALL_CASES as CASES, ATTEN, FIELD_DIMS, E, L2, load, state
from tests.test_oracle_golden import bias_before_bn, close


class Cfg:
    use_atten = False
    use_dcn = False
    dataset_name = "synthetic"
    mmoe_n_expert = 3
    ple_n_expert_specific = 2
    ple_n_expert_shared = 1


KIND_CLS = {"ple": "PLE", "mmoe": "MMoE", "dcn": "DCN", "dcnv2": "DCNv2", "star": "STAR", "autoint": "AutoInt"}
DOMAIN_IDX = 3


def build_model(name, probe=False, precision="fp32"):
    kind, kw, mode, steps = CASES[name]
    is_cdc = name.startswith("cdc_")
    if not hasattr(cm, KIND_CLS[kind]) or (is_cdc and not hasattr(cm, "CDC")):
        if probe:
            return None
        raise NotImplementedError(name)
    if probe:
        return True
    cfg = Cfg()
    cfg.cdcmdr_precision = precision
    if name in ATTEN:
        cfg.use_atten = True
        for k, v in ATTEN[name].items():
            setattr(cfg, k, v)
    if is_cdc:
        gold = load(name)
        n_domain = len(gold["d2g"])
        kw2 = dict(kw)
        T = kw2.pop("n_tower")
        base_kw = dict(expert_dims=kw2.get("expert_dims"), tower_dims=kw2["tower_dims"])
        model = cm.CDC(FIELD_DIMS, E, T, n_domain, kind, base_kw["expert_dims"], base_kw["tower_dims"], DOMAIN_IDX,
                       domain_cnt_weight=[1.0 / n_domain] * n_domain, dropout=0.0, config=cfg, **L2)
        model.set_groups(gold["d2g"].tolist())
        return model
    cls = getattr(cm, KIND_CLS[kind])
    extra = {}
    if kind in ("ple", "mmoe", "star"):
        extra["config"] = cfg
    if kind == "star":
        extra["domain_idx"] = DOMAIN_IDX
    if kind == "autoint":
        return cls(FIELD_DIMS, E, dropout=0.0, config=cfg, **kw, **{k: v for k, v in L2.items() if k != "l2_reg_cross"})
    return cls(FIELD_DIMS, E, dropout=0.0, **kw, **extra, **L2)


def _t(a, device):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def run_golden_case(name, device, path="fused", rtol=1e-4, atol=2e-6, precision="fp32", loose=1.0):
    kind, kw, mode, steps = CASES[name]
    gold = load(name)
    model = build_model(name, precision=precision)
    is_cdc = name.startswith("cdc_")
    prefix = "base_model_instance." if is_cdc else ""
    sd0 = {prefix + k: torch.from_numpy(v) for k, v in state(gold, 0).items()}
    model.load_state_dict(sd0, strict=True)
    model = model.to(device)
    base = model.base_model_instance if is_cdc else model
    opt = cm.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)
    opt.attach(base)
    crit = torch.nn.BCELoss()
    model.train()
    for s in range(steps):
        x, y, g = _t(gold[f"in{s}.x"], device), _t(gold[f"in{s}.y"], device), _t(gold[f"in{s}.g"], device)
        if path == "fused":
            if is_cdc:
                cmode = "warmup" if mode == "warmup" else "split"
                di = int(gold["domain_i"]) if mode == "split_domain" else None
                out = model.train_step(x, y, opt, mode=cmode, domain_i=di)
            elif mode == "gather":
                out = model.train_step(x, y, opt, mode="gather", sel=g)
            elif mode == "single":
                out = model.train_step(x, y, opt, mode="col", col=0)
            elif mode == "star_grouped":
                out = model.train_step(x, y, opt, mode="col", col=0, x_group=g)
            else:
                raise NotImplementedError(mode)
            loss, bce, reg = base.step_losses(out)
            pred = out["pred"].detach().cpu().numpy()
            psel = out["psel"].detach().cpu().numpy()
        else:
            if is_cdc:
                cmode = "warmup" if mode == "warmup" else "split"
                di = int(gold["domain_i"]) if mode == "split_domain" else None
                p_sel = model(x, mode=cmode, domain_i=di).reshape(-1)
                tgt = y.float().reshape(-1)
                pred = None
            elif mode == "gather":
                full = model(x)
                p_sel, tgt = full.gather(1, g).squeeze(1), y.float().squeeze(1)
                pred = full.detach().cpu().numpy()
            elif mode == "single":
                full = model(x)
                p_sel, tgt = full.reshape(-1), y.float().reshape(-1)
                pred = full.detach().cpu().numpy()
            elif mode == "star_grouped":
                full, tperm = model(x, g, targets=y)
                p_sel, tgt = full.reshape(-1), tperm.float().reshape(-1)
                pred = full.detach().cpu().numpy()
            bce_t = crit(p_sel, tgt)
            reg_t = model.get_regularization_loss(device=device)
            loss_t = bce_t + reg_t
            model.zero_grad()
            loss_t.backward()
            if s == 0:
                gk = {k[6:]: v for k, v in gold.items() if k.startswith("grad0.")}
                named = dict(model.named_parameters())
                got = {k for k, p in named.items() if p.grad is not None}
                assert got == set(gk), got ^ set(gk)
                for k, v in gk.items():
                    kk = k[len(prefix):] if prefix and k.startswith(prefix) else k
                    # gradients: 1e-4 relative to the TENSOR's scale (summation order differs from torch's, so entries
                    # that cancel to ~0 carry rounding noise of the order of eps * the largest entries)
                    a = (1e-4 if bias_before_bn(kind, kk) else (1e-6 if kk.endswith('.bias') else 2e-7)) * loose
                    close(named[k].grad.detach().cpu().numpy().reshape(v.shape), v, f"{name} grad {k}", rtol=rtol * loose,
                          atol=a + rtol * loose * float(np.abs(v).max()))
            opt.step()
            bce, reg = float(bce_t.detach()), float(reg_t.detach())
            psel = p_sel.detach().cpu().numpy()
        gp = gold[f"step{s}.pred"]
        if mode == "star_grouped":
            close(pred.reshape(-1), gp[:, 0], f"{name} step{s} pred", rtol * loose, atol * loose)
        elif is_cdc:
            close(psel.reshape(-1), gp.reshape(-1), f"{name} step{s} pred", rtol * loose, atol * loose)
        elif pred is not None:
            close(pred.reshape(gp.shape), gp, f"{name} step{s} pred", rtol * loose, atol * loose)
        close(bce, gold[f"step{s}.bce"], "bce", rtol * loose, atol * loose)
        close(reg, gold[f"step{s}.reg"], "reg", rtol * loose, atol * loose)
        ref = state(gold, s + 1)
        cur = {k[len(prefix):] if prefix else k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
        assert set(ref) == set(cur), set(ref) ^ set(cur)
        for k, v in ref.items():
            a = (2.1e-3 * (s + 1) if bias_before_bn(kind, k) else (3e-4 * (s + 1) if k.endswith("running_mean") else 1e-6))
            if k.endswith("in_proj_bias"):
                # the KEY bias of an attention layer shifts every score of a row by the same q.b_k: softmax cancels it, its gradient
                # is rounding noise, and Adam turns the noise's sign into +-lr steps (in the reference as much as here) - the same
                # allowance as for a bias in front of a batch-statistics BatchNorm; the query / value thirds are checked tightly
                n3 = v.shape[0] // 3
                c = cur[k].reshape(v.shape)
                close(c[n3:2 * n3], v[n3:2 * n3], f"{name} step{s + 1} {k} (key third)", rtol * loose, 2.1e-3 * (s + 1) * loose)
                close(np.delete(c, np.s_[n3:2 * n3]), np.delete(v, np.s_[n3:2 * n3]), f"{name} step{s + 1} {k}", rtol * loose, a * loose)
                continue
            close(cur[k].reshape(v.shape), v, f"{name} step{s + 1} {k}", rtol * loose, a * loose)
    # eval-mode forward with the reference's final weights
    model.load_state_dict({prefix + k: torch.from_numpy(v) for k, v in state(gold, steps).items()}, strict=True)
    model.eval()
    x, g = _t(gold["in0.x"], device), _t(gold["in0.g"], device)
    with torch.no_grad():
        if is_cdc:
            cmode = "warmup" if mode == "warmup" else "split"
            di = int(gold["domain_i"]) if mode == "split_domain" else None
            p = model(x, mode=cmode, domain_i=di)
        elif mode == "star_grouped":
            p = model(x, g)
        else:
            p = model(x)
    ge = gold["eval.pred"]
    p = p.detach().cpu().numpy()
    if mode == "star_grouped":
        close(p.reshape(-1), ge[:, 0], "eval pred", rtol * loose, atol * loose)
    else:
        close(p.reshape(-1), ge.reshape(-1), "eval pred", rtol * loose, atol * loose)
    return model
