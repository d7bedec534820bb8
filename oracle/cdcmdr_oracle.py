"""CPU oracle for the CDC-MDR multi-domain CTR hot path (TEST INFRASTRUCTURE, not product).

A plain-numpy restatement of the reference's algorithm for the path SURVEY.md §8(a) names:
embedding gather, PLE/CGC, MMoE, DCN / DCNv2 cross networks, STAR, CDC tower selection,
BCE + L2 regularisation, and the explicit backward pass and dense Adam step that
`loss.backward(); optimizer.step()` perform in the reference (run.py:489-492, 720-721).
Every function cites the reference file:line it follows (paths relative to /root/reference).

PARITY PIN: the reference ships no tests or golden vectors.  This oracle is
pinned against fixtures produced by importing and running the unmodified reference in the
dev container (tests/golden/make_golden.py -> tests/golden/*.npz); tests/test_oracle_golden.py
checks forward, loss, every gradient, BN running stats and post-Adam parameters over several steps.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package never does.

All tensors are numpy arrays keyed by the reference's state_dict names.  The arithmetic runs in the
dtype of the state dict handed in: float32 reproduces the reference's precision (what the golden fixtures pin);
float64 gives the exact value of the same algorithm, which the GPU parity tests use at sizes where float32 numpy
itself is noisy (a BatchNorm unit that is constant over the batch: rounding-noise xhat, coin-flip ReLU mask).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
BN_EPS = 1e-5
BN_MOM = 0.1


# --------------------------------------------------------------------------------------
# primitive ops
# --------------------------------------------------------------------------------------
def field_offsets(field_dims):
    """model/layer.py:141-144: offsets = (0, cumsum(field_dims)[:-1])."""
    fd = np.asarray(field_dims, dtype=np.int64)
    return np.concatenate([[0], np.cumsum(fd)[:-1]]).astype(np.int64)


def embed_gather(x, offsets, table):
    """model/layer.py:147-157 with squeeze_dim=True: out[b, f*E:(f+1)*E] = table[x[b,f]+offsets[f]]."""
    idx = x.astype(np.int64) + offsets[None, :]
    return table[idx].reshape(x.shape[0], -1), idx


def embed_scatter_grad(idx, dout, V, E):
    """autograd of nn.Embedding(sparse=False) (model/layer.py:140): dense [V,E] scatter-add.

    Summation order inside a row is ascending (b, f) position - the deterministic order the
    CUDA sorted-segment kernel uses."""
    g = np.zeros((V, E), dtype=dout.dtype)
    flat = idx.reshape(-1)
    order = np.argsort(flat, kind="stable")
    rows = flat[order]
    vals = dout.reshape(-1, E)[order]
    if rows.size:
        starts = np.flatnonzero(np.concatenate([[True], rows[1:] != rows[:-1]]))
        g[rows[starts]] = np.add.reduceat(vals, starts, axis=0)
    return g


def linear_fwd(x, W, b=None):
    y = x @ W.T
    if b is not None:
        y = y + b
    return y


def linear_bwd(x, W, dy, has_bias=True):
    dx = dy @ W
    dW = dy.T @ x
    db = dy.sum(0) if has_bias else None
    return dx, dW, db


def softmax_rows(z):
    z = z - z.max(axis=1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=1, keepdims=True)


def softmax_bwd(p, dp):
    return p * (dp - (dp * p).sum(axis=1, keepdims=True))


def sigmoid(z):
    return (1.0 / (1.0 + np.exp(-z))).astype(z.dtype)


def bn_fwd(z, gamma, beta, rmean, rvar, train):
    """torch BatchNorm1d / F.batch_norm: train -> biased batch var for the
    normalisation, running_var updated with the UNBIASED var, momentum 0.1, eps 1e-5."""
    if train:
        n = z.shape[0]
        mu = z.mean(0)
        var = z.var(0)
        inv = 1.0 / np.sqrt(var + BN_EPS)
        xhat = (z - mu) * inv
        new_rm = (1 - BN_MOM) * rmean + BN_MOM * mu
        new_rv = (1 - BN_MOM) * rvar + BN_MOM * var * (n / max(n - 1, 1))
        return (xhat * gamma + beta).astype(z.dtype), (xhat, inv, gamma), (new_rm.astype(rmean.dtype), new_rv.astype(rvar.dtype))
    inv = 1.0 / np.sqrt(rvar + BN_EPS)
    xhat = (z - rmean) * inv
    return (xhat * gamma + beta).astype(z.dtype), (xhat, inv, gamma), None


def bn_bwd(dy, cache, train):
    xhat, inv, gamma = cache
    dgamma = (dy * xhat).sum(0)
    dbeta = dy.sum(0)
    if train:
        n = dy.shape[0]
        dz = (gamma * inv / n) * (n * dy - dbeta - xhat * dgamma)
    else:
        dz = dy * gamma * inv
    return dz.astype(dy.dtype), dgamma, dbeta


def bce_mean(p, y):
    """torch BCELoss(mean) (run.py:723): logs clamped at -100; backward divides by
    max(p(1-p), 1e-12)."""
    p64 = p.astype(np.float64)
    lp = np.maximum(np.log(p64), -100.0)
    l1p = np.maximum(np.log1p(-p64), -100.0)
    loss = -(y * lp + (1 - y) * l1p).mean()
    dp = (p64 - y) / np.maximum((1 - p64) * p64, 1e-12) / p.shape[0]
    return F32(loss), dp.astype(p.dtype)


# --------------------------------------------------------------------------------------
# MultiLayerPerceptron / DNN (model/layer.py:178-206, 238-300)
# --------------------------------------------------------------------------------------
class MLP:
    """Reference MultiLayerPerceptron: per layer Linear -> [BN] -> ReLU -> Dropout, optional
    final Linear(.,1).  Parameter keys `<prefix>.layers.N.*` (stride 4 with BN, 3 without)."""

    def __init__(self, prefix, n_hidden, bn, output_layer):
        self.prefix, self.n_hidden, self.bn, self.out = prefix, n_hidden, bn, output_layer
        s = 4 if bn else 3
        self.lin = [f"{prefix}.layers.{s * j}" for j in range(n_hidden)]
        self.bnk = [f"{prefix}.layers.{s * j + 1}" for j in range(n_hidden)] if bn else []
        self.outk = f"{prefix}.layers.{s * n_hidden}" if output_layer else None

    def forward(self, sd, x, train, bufs, drop=None):
        cache = []
        h = x
        use_bn = self.bn and x.shape[0] != 1          # layer.py:202-204
        for j in range(self.n_hidden):
            z = linear_fwd(h, sd[self.lin[j] + ".weight"], sd[self.lin[j] + ".bias"])
            bc = None
            if use_bn:
                k = self.bnk[j]
                z, bc, upd = bn_fwd(z, sd[k + ".weight"], sd[k + ".bias"], sd[k + ".running_mean"],
                                    sd[k + ".running_var"], train)
                if upd is not None:
                    bufs[k + ".running_mean"], bufs[k + ".running_var"] = upd
                    bufs[k + ".num_batches_tracked"] = sd[k + ".num_batches_tracked"] + 1
            a = np.maximum(z, 0)
            m = None
            if drop is not None and train:
                m = drop(a.shape)
                a = a * m
            cache.append((h, bc, a > 0 if m is None else (z > 0), m))
            h = a
        if self.out:
            z = linear_fwd(h, sd[self.outk + ".weight"], sd[self.outk + ".bias"])
            cache.append((h,))
            h = z
        return h, (cache, use_bn, train)

    def backward(self, sd, cache, dy, grads):
        cache, use_bn, train = cache
        if self.out:
            (h,) = cache[-1]
            dy, dW, db = linear_bwd(h, sd[self.outk + ".weight"], dy)
            _acc(grads, self.outk + ".weight", dW); _acc(grads, self.outk + ".bias", db)
        for j in reversed(range(self.n_hidden)):
            h, bc, mask, m = cache[j]
            if m is not None:
                dy = dy * m
            dz = dy * mask
            if use_bn:
                dz, dg, dbt = bn_bwd(dz, bc, train)
                _acc(grads, self.bnk[j] + ".weight", dg); _acc(grads, self.bnk[j] + ".bias", dbt)
            dy, dW, db = linear_bwd(h, sd[self.lin[j] + ".weight"], dz)
            _acc(grads, self.lin[j] + ".weight", dW); _acc(grads, self.lin[j] + ".bias", db)
        return dy


def _acc(grads, k, g):
    g = np.asarray(g)
    grads[k] = g.copy() if k not in grads else grads[k] + g


# --------------------------------------------------------------------------------------
# Base: embedding + linear + regularisation bookkeeping (model/layer.py:10-112)
# --------------------------------------------------------------------------------------
class Base:
    reg_prefixes = ()          # [(prefix, l2)] sub-modules registered with the 'weight' & not 'bn' filter
    reg_exact = ()             # [(key, l2)] explicit parameters

    def __init__(self, field_dims, embed_dim, l2_reg_embedding=1e-5, l2_reg_linear=1e-5):
        self.field_dims = np.asarray(field_dims, dtype=np.int64)
        self.offsets = field_offsets(field_dims)
        self.F, self.E = len(self.field_dims), embed_dim
        self.D = self.F * embed_dim
        self.V = int(self.field_dims.sum())
        self.l2_emb, self.l2_lin = l2_reg_embedding, l2_reg_linear
        self.drop = None
        self.atten = None          # enable_atten(): the field self-attention block of BaseModel.build_atten (layer.py:58-69)
        self._att_cache = None

    # ---------------------------------------------------------------- field self-attention (model/layer.py:58-84)
    def enable_atten(self, atten_embed_dim, att_layer_num, att_head_num, att_res=True, head=("atten_linear.weight", 0)):
        """config.use_atten with config.atten_embed_dim / att_layer_num / att_head_num / att_res (config.py:24-28).  The block is
        not regularised (none of the models registers its parameters, ple.py:42-48).  head: (key, first column) of the bias-free
        Linear over relu(block output).view(B, F*A) - `atten_linear` for BaseModel.atten_forward, the first F*A columns of
        `dnn_linear` for AutoInt (autoint.py:60-62)."""
        self.atten = dict(A=int(atten_embed_dim), n_layer=int(att_layer_num), H=int(att_head_num), res=bool(att_res), head=head)
        return self

    def atten_fwd(self, sd, e):
        """atten_forward (layer.py:71-84): tokens = embed_x.view(B, F, E); atten_embedding; att_layer_num x nn.MultiheadAttention
        (in_proj -> per-head softmax(q k^T / sqrt(dh)) v -> out_proj; no residual, no norm between layers; attention dropout is
        inactive at dropout = 0 / eval, the only settings the oracle is used at); + V_res_embedding(tokens) if att_res; ReLU;
        view(B, F*A); atten_linear (no bias) -> (B, 1)."""
        a = self.atten
        B, F, E, A, H = e.shape[0], self.F, self.E, a["A"], a["H"]
        dh = A // H
        tok = e.reshape(B * F, E)
        cur = linear_fwd(tok, sd["atten_embedding.weight"], sd["atten_embedding.bias"])
        layers = []
        for i in range(a["n_layer"]):
            pre = f"self_attns.{i}."
            qkv = linear_fwd(cur, sd[pre + "in_proj_weight"], sd[pre + "in_proj_bias"])
            q, k, v = (qkv[:, j * A:(j + 1) * A].reshape(B, F, H, dh).transpose(0, 2, 1, 3) for j in range(3))     # [B, H, F, dh]
            sc = np.einsum("bhid,bhjd->bhij", q, k).astype(F32) * F32(1.0 / np.sqrt(dh))
            sc = sc - sc.max(axis=-1, keepdims=True)
            p = np.exp(sc).astype(F32)
            p = (p / p.sum(axis=-1, keepdims=True, dtype=F32)).astype(F32)
            o = np.einsum("bhij,bhjd->bhid", p, v).astype(F32).transpose(0, 2, 1, 3).reshape(B * F, A)
            layers.append(dict(x=cur, q=q, k=k, v=v, p=p, o=o))
            cur = linear_fwd(o, sd[pre + "out_proj.weight"], sd[pre + "out_proj.bias"])
        if a["res"]:
            cur = cur + linear_fwd(tok, sd["V_res_embedding.weight"], sd["V_res_embedding.bias"])
        r = np.maximum(cur, 0).reshape(B, F * A)
        hk, h0 = a["head"]
        out = linear_fwd(r, sd[hk][:, h0:h0 + F * A])
        return out, dict(tok=tok, layers=layers, z=cur, r=r)

    def atten_bwd(self, sd, c, dout, grads):
        """gradient of atten_fwd's (B, 1) output: parameter gradients into `grads`, returns d(embed_x) [B, F*E]."""
        a = self.atten
        A, H = a["A"], a["H"]
        dh = A // H
        tok = c["tok"]
        M = tok.shape[0]
        B = M // self.F
        hk, h0 = a["head"]
        dr, dW, _ = linear_bwd(c["r"], sd[hk][:, h0:h0 + self.F * A], dout, has_bias=False)
        dfull = np.zeros_like(sd[hk])
        dfull[:, h0:h0 + self.F * A] = dW
        _acc(grads, hk, dfull)
        dy = (dr.reshape(M, A) * (c["z"] > 0)).astype(F32)
        dtok = np.zeros_like(tok)
        if a["res"]:
            dx, dW, db = linear_bwd(tok, sd["V_res_embedding.weight"], dy)
            _acc(grads, "V_res_embedding.weight", dW); _acc(grads, "V_res_embedding.bias", db)
            dtok += dx
        for i in reversed(range(a["n_layer"])):
            pre, L = f"self_attns.{i}.", c["layers"][i]
            do, dW, db = linear_bwd(L["o"], sd[pre + "out_proj.weight"], dy)
            _acc(grads, pre + "out_proj.weight", dW); _acc(grads, pre + "out_proj.bias", db)
            do = do.reshape(B, self.F, H, dh).transpose(0, 2, 1, 3)
            dv = np.einsum("bhij,bhid->bhjd", L["p"], do)
            dp = np.einsum("bhid,bhjd->bhij", do, L["v"])
            ds = (L["p"] * (dp - (L["p"] * dp).sum(axis=-1, keepdims=True)) * F32(1.0 / np.sqrt(dh))).astype(F32)
            dq = np.einsum("bhij,bhjd->bhid", ds, L["k"])
            dk = np.einsum("bhij,bhid->bhjd", ds, L["q"])
            dqkv = np.concatenate([t.astype(F32).transpose(0, 2, 1, 3).reshape(M, A) for t in (dq, dk, dv)], axis=1)
            dy, dW, db = linear_bwd(L["x"], sd[pre + "in_proj_weight"], dqkv)
            _acc(grads, pre + "in_proj_weight", dW); _acc(grads, pre + "in_proj_bias", db)
        dx, dW, db = linear_bwd(tok, sd["atten_embedding.weight"], dy)
        _acc(grads, "atten_embedding.weight", dW); _acc(grads, "atten_embedding.bias", db)
        dtok += dx
        return dtok.reshape(B, self.F * self.E)

    # model/layer.py:31-33, 86-112 and the per-model add_regularization_weight calls
    def reg_items(self, sd):
        items = [("embedding.embedding_dict.weight", self.l2_emb), ("linear.fc.weight", self.l2_lin)]
        for prefix, l2 in self.reg_prefixes:
            for k in sd:
                if k.startswith(prefix + "."):
                    rel = k[len(prefix) + 1:]
                    if "weight" in rel and "bn" not in rel and _is_param(k):
                        items.append((k, l2))
        items += list(self.reg_exact)
        return items

    def reg_loss(self, sd):
        tot = np.zeros((), dtype=np.float64)
        for k, l2 in self.reg_items(sd):
            if l2 > 0:
                tot += np.float64(F32(l2)) * np.square(sd[k].astype(np.float64)).sum()
        return F32(tot)

    def add_reg_grad(self, sd, grads):
        for k, l2 in self.reg_items(sd):
            if l2 > 0:
                _acc(grads, k, (2.0 * l2) * sd[k])

    def embed(self, sd, x):
        return embed_gather(x, self.offsets, sd["embedding.embedding_dict.weight"])

    def lin_fwd(self, sd, e):
        """The `other_outs` every tower logit receives (layer.py:52-54): FeaturesLinear, plus the field self-attention scalar when
        the model was built with config.use_atten (ple.py:65-67, mmoe.py:68-70, star.py:70-72)."""
        lin = linear_fwd(e, sd["linear.fc.weight"], sd["linear.fc.bias"])
        self._att_cache = None
        if self.atten is not None:
            a, self._att_cache = self.atten_fwd(sd, e)
            lin = lin + a
        return lin

    def embed_bwd(self, sd, idx, e, dembed, dlin, grads):
        """FeaturesLinear backward (layer.py:122-126) + embedding scatter (layer.py:153)."""
        if dlin is not None:
            dx, dW, db = linear_bwd(e, sd["linear.fc.weight"], dlin)
            _acc(grads, "linear.fc.weight", dW); _acc(grads, "linear.fc.bias", db)
            dembed = dembed + dx
            if self._att_cache is not None:
                dembed = dembed + self.atten_bwd(sd, self._att_cache, dlin, grads)
        _acc(grads, "embedding.embedding_dict.weight", embed_scatter_grad(idx, dembed, self.V, self.E))

    def towers_fwd(self, sd, tower_in, lin, train, bufs):
        """layer.py:48-56: logit_t = tower_t(in_t) + lin ; y_t = sigmoid ; cat."""
        ys, caches = [], []
        for t, mlp in enumerate(self.towers):
            z, c = mlp.forward(sd, tower_in[t], train, bufs, self.drop)
            ys.append(sigmoid(z + lin))
            caches.append(c)
        return np.concatenate(ys, axis=1), caches

    def towers_bwd(self, sd, caches, y, dy, grads):
        dlogit = dy * y * (1 - y)
        dins = [mlp.backward(sd, caches[t], dlogit[:, t:t + 1], grads) for t, mlp in enumerate(self.towers)]
        return dins, dlogit.sum(1, keepdims=True)


def _is_param(k):
    return not (k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"))


# --------------------------------------------------------------------------------------
# PLE / CGC (model/ple.py:9-124)
# --------------------------------------------------------------------------------------
class PLE(Base):
    def __init__(self, field_dims, embed_dim, n_tower, n_expert_specific, n_expert_shared, expert_dims, tower_dims,
                 l2_reg_embedding=1e-5, l2_reg_linear=1e-5, l2_reg_dnn=1e-5, **_):
        super().__init__(field_dims, embed_dim, l2_reg_embedding, l2_reg_linear)
        self.T, self.ns, self.nsh = n_tower, n_expert_specific, n_expert_shared
        self.n_level = len(expert_dims)
        self.levels = []
        for l, dims in enumerate(expert_dims):
            p = f"cgc_layers.{l}"
            self.levels.append(dict(
                spec=[MLP(f"{p}.experts_specific.{i}", len(dims), False, False) for i in range(self.T * self.ns)],
                shared=[MLP(f"{p}.experts_shared.{i}", len(dims), False, False) for i in range(self.nsh)],
                gates=[f"{p}.gates_specific.{t}.0" for t in range(self.T)],
                gate_shared=f"{p}.gate_shared.0" if l + 1 < self.n_level else None))
        self.towers = [MLP(f"towers.{t}", len(tower_dims), True, True) for t in range(self.T)]
        self.reg_prefixes = (("cgc_layers", l2_reg_dnn), ("towers", l2_reg_dnn))

    def _cgc_fwd(self, sd, lv, xs, train, bufs):
        """ple.py:96-124.  Task gate t mixes [own specific..., shared...]; the shared gate (non-last
        level) mixes [all specific task-major..., shared...]."""
        T, ns = self.T, self.ns
        eo, ec = [], []
        for i, mlp in enumerate(lv["spec"]):
            o, c = mlp.forward(sd, xs[i // ns], train, bufs, self.drop); eo.append(o); ec.append(c)
        for mlp in lv["shared"]:
            o, c = mlp.forward(sd, xs[-1], train, bufs, self.drop); eo.append(o); ec.append(c)
        outs, gs = [], []
        for t in range(T):
            g = softmax_rows(linear_fwd(xs[t], sd[lv["gates"][t] + ".weight"], sd[lv["gates"][t] + ".bias"]))
            sel = list(range(t * ns, (t + 1) * ns)) + list(range(T * ns, T * ns + self.nsh))
            outs.append(sum(g[:, j:j + 1] * eo[e] for j, e in enumerate(sel)))
            gs.append((g, sel))
        if lv["gate_shared"] is not None:
            g = softmax_rows(linear_fwd(xs[-1], sd[lv["gate_shared"] + ".weight"], sd[lv["gate_shared"] + ".bias"]))
            sel = list(range(T * ns + self.nsh))
            outs.append(sum(g[:, j:j + 1] * eo[e] for j, e in enumerate(sel)))
            gs.append((g, sel))
        return outs, (xs, eo, ec, gs)

    def _cgc_bwd(self, sd, lv, cache, douts, grads):
        xs, eo, ec, gs = cache
        T, ns = self.T, self.ns
        dxs = [np.zeros_like(xs[i]) for i in range(len(xs))]
        deo = [np.zeros_like(o) for o in eo]
        for gi, (g, sel) in enumerate(gs):
            do = douts[gi]
            src = gi if gi < T else len(xs) - 1
            key = lv["gates"][gi] if gi < T else lv["gate_shared"]
            dg = np.stack([(do * eo[e]).sum(1) for e in sel], axis=1)
            for j, e in enumerate(sel):
                deo[e] += g[:, j:j + 1] * do
            dz = softmax_bwd(g, dg)
            dx, dW, db = linear_bwd(xs[src], sd[key + ".weight"], dz)
            _acc(grads, key + ".weight", dW); _acc(grads, key + ".bias", db)
            dxs[src] = dxs[src] + dx
        for i, mlp in enumerate(lv["spec"]):
            dxs[i // ns] = dxs[i // ns] + mlp.backward(sd, ec[i], deo[i], grads)
        for i, mlp in enumerate(lv["shared"]):
            dxs[-1] = dxs[-1] + mlp.backward(sd, ec[T * ns + i], deo[T * ns + i], grads)
        return dxs

    def forward(self, sd, x, train=True):
        """ple.py:50-70."""
        bufs = {}
        e, idx = self.embed(sd, x)
        xs = [e] * (self.T + 1)
        lc = []
        for lv in self.levels:
            xs, c = self._cgc_fwd(sd, lv, xs, train, bufs); lc.append(c)
        lin = self.lin_fwd(sd, e)
        y, tc = self.towers_fwd(sd, xs[:self.T], lin, train, bufs)
        return y, dict(idx=idx, e=e, lc=lc, tc=tc, y=y, bufs=bufs)

    def backward(self, sd, cache, dy):
        grads = {}
        dins, dlin = self.towers_bwd(sd, cache["tc"], cache["y"], dy, grads)
        douts = dins
        for l in reversed(range(self.n_level)):
            nx = len(cache["lc"][l][0])
            if len(douts) < len(cache["lc"][l][3]):      # never: last level has exactly T outputs
                raise AssertionError
            douts = self._cgc_bwd(sd, self.levels[l], cache["lc"][l], douts, grads)
            assert len(douts) == nx
        dembed = sum(douts)                               # level 0 inputs are all embed_x (ple.py:54)
        self.embed_bwd(sd, cache["idx"], cache["e"], dembed, dlin, grads)
        return grads


# --------------------------------------------------------------------------------------
# MMoE (model/mmoe.py:10-74)
# --------------------------------------------------------------------------------------
class MMoE(Base):
    def __init__(self, field_dims, embed_dim, n_tower, n_expert, expert_dims, tower_dims,
                 l2_reg_embedding=1e-5, l2_reg_linear=1e-5, l2_reg_dnn=1e-5, **_):
        super().__init__(field_dims, embed_dim, l2_reg_embedding, l2_reg_linear)
        self.T, self.ne = n_tower, n_expert
        self.experts = [MLP(f"experts.{i}", len(expert_dims), True, False) for i in range(n_expert)]
        self.towers = [MLP(f"towers.{t}", len(tower_dims), True, True) for t in range(n_tower)]
        self.reg_prefixes = (("experts", l2_reg_dnn), ("towers", l2_reg_dnn))

    def forward(self, sd, x, train=True):
        """mmoe.py:53-74: tower_in_t = sum_e softmax(gate_t(embed))_e * expert_e(embed)."""
        bufs = {}
        e, idx = self.embed(sd, x)
        eo, ec = [], []
        for mlp in self.experts:
            o, c = mlp.forward(sd, e, train, bufs, self.drop); eo.append(o); ec.append(c)
        gs, tin = [], []
        for t in range(self.T):
            g = softmax_rows(linear_fwd(e, sd[f"gates.{t}.0.weight"], sd[f"gates.{t}.0.bias"]))
            gs.append(g)
            tin.append(sum(g[:, j:j + 1] * eo[j] for j in range(self.ne)))
        lin = self.lin_fwd(sd, e)
        y, tc = self.towers_fwd(sd, tin, lin, train, bufs)
        return y, dict(idx=idx, e=e, eo=eo, ec=ec, gs=gs, tc=tc, y=y, bufs=bufs)

    def backward(self, sd, cache, dy):
        grads = {}
        e, eo = cache["e"], cache["eo"]
        dins, dlin = self.towers_bwd(sd, cache["tc"], cache["y"], dy, grads)
        dembed = np.zeros_like(e)
        deo = [np.zeros_like(o) for o in eo]
        for t in range(self.T):
            g = cache["gs"][t]
            dg = np.stack([(dins[t] * eo[j]).sum(1) for j in range(self.ne)], axis=1)
            for j in range(self.ne):
                deo[j] += g[:, j:j + 1] * dins[t]
            dx, dW, db = linear_bwd(e, sd[f"gates.{t}.0.weight"], softmax_bwd(g, dg))
            _acc(grads, f"gates.{t}.0.weight", dW); _acc(grads, f"gates.{t}.0.bias", db)
            dembed += dx
        for j, mlp in enumerate(self.experts):
            dembed += mlp.backward(sd, cache["ec"][j], deo[j], grads)
        self.embed_bwd(sd, cache["idx"], e, dembed, dlin, grads)
        return grads


# --------------------------------------------------------------------------------------
# Cross networks (model/layer.py:495-515, 332-343, 346-407)
# --------------------------------------------------------------------------------------
def cross_v1_fwd(sd, p, L, x0):
    """layer.py:495-515 (second, live definition): x <- x0 * (x . w_l) + b_l + x."""
    xs = [x0]
    x = x0
    for l in range(L):
        xw = x @ sd[f"{p}.w.{l}.weight"].T               # (B,1)
        x = x0 * xw + sd[f"{p}.b.{l}"] + x
        xs.append(x)
    return x, xs


def cross_v1_bwd(sd, p, L, xs, dout, grads):
    x0 = xs[0]
    dx = dout
    dx0 = np.zeros_like(x0)
    for l in reversed(range(L)):
        w = sd[f"{p}.w.{l}.weight"]
        xw = xs[l] @ w.T
        _acc(grads, f"{p}.b.{l}", dx.sum(0))
        dxw = (dx * x0).sum(1, keepdims=True)
        dx0 += dx * xw
        _acc(grads, f"{p}.w.{l}.weight", dxw.T @ xs[l])
        dx = dx + dxw @ w
    return dx + dx0


def cross_v2_fwd(sd, p, L, x0):
    """layer.py:332-343: x <- x0 * (x W_l^T) + b_l + x   (bias OUTSIDE the product)."""
    xs = [x0]
    x = x0
    for l in range(L):
        x = x0 * (x @ sd[f"{p}.w.{l}.weight"].T) + sd[f"{p}.b.{l}"] + x
        xs.append(x)
    return x, xs


def cross_v2_bwd(sd, p, L, xs, dout, grads):
    x0 = xs[0]
    dx = dout
    dx0 = np.zeros_like(x0)
    for l in reversed(range(L)):
        W = sd[f"{p}.w.{l}.weight"]
        xw = xs[l] @ W.T
        _acc(grads, f"{p}.b.{l}", dx.sum(0))
        dxw = dx * x0
        dx0 += dx * xw
        _acc(grads, f"{p}.w.{l}.weight", dxw.T @ xs[l])
        dx = dx + dxw @ W
    return dx + dx0


def cross_mix_fwd(sd, p, L, n_exp, x0):
    """layer.py:380-407.  Per layer / expert e: v = tanh(V_e^T x); v = tanh(C_e v); u = U_e v;
    out_e = x0 * (u + bias_l); x <- x + sum_e softmax_e(gating_e(x)) * out_e.  gating is shared
    by all layers."""
    G = np.concatenate([sd[f"{p}.gating.{e}.weight"] for e in range(n_exp)], axis=0)    # (n_exp, D)
    x = x0
    cache = []
    for l in range(L):
        U, Vm, C = sd[f"{p}.u_list.{l}"], sd[f"{p}.v_list.{l}"], sd[f"{p}.c_list.{l}"]
        bias = sd[f"{p}.bias.{l}"][:, 0]
        g = softmax_rows(x @ G.T)
        v1 = np.tanh(np.einsum("bd,edr->ebr", x, Vm))
        v2 = np.tanh(np.einsum("ers,ebs->ebr", C, v1))
        u = np.einsum("edr,ebr->ebd", U, v2)
        out = x0[None] * (u + bias)
        xn = x + np.einsum("ebd,be->bd", out, g)
        cache.append((x, g, v1, v2, u, out))
        x = xn.astype(x0.dtype)
    return x, cache


def cross_mix_bwd(sd, p, L, n_exp, x0, cache, dout, grads):
    G = np.concatenate([sd[f"{p}.gating.{e}.weight"] for e in range(n_exp)], axis=0)
    dG = np.zeros_like(G)
    dx = dout
    dx0 = np.zeros_like(x0)
    for l in reversed(range(L)):
        U, Vm, C = sd[f"{p}.u_list.{l}"], sd[f"{p}.v_list.{l}"], sd[f"{p}.c_list.{l}"]
        bias = sd[f"{p}.bias.{l}"][:, 0]
        x, g, v1, v2, u, out = cache[l]
        dgate = np.einsum("bd,ebd->be", dx, out)
        dout_e = dx[None] * g.T[:, :, None]                      # (e,B,D)
        dx0 += (dout_e * (u + bias)).sum(0)
        du = dout_e * x0[None]
        _acc(grads, f"{p}.bias.{l}", du.sum((0, 1))[:, None])
        _acc(grads, f"{p}.u_list.{l}", np.einsum("ebd,ebr->edr", du, v2))
        dv2 = np.einsum("ebd,edr->ebr", du, U) * (1 - v2 * v2)
        _acc(grads, f"{p}.c_list.{l}", np.einsum("ebr,ebs->ers", dv2, v1))
        dv1 = np.einsum("ebr,ers->ebs", dv2, C) * (1 - v1 * v1)
        _acc(grads, f"{p}.v_list.{l}", np.einsum("bd,ebr->edr", x, dv1))
        dz = softmax_bwd(g, dgate)
        dG += dz.T @ x
        dx = dx + np.einsum("ebr,edr->bd", dv1, Vm) + dz @ G
    for e in range(n_exp):
        _acc(grads, f"{p}.gating.{e}.weight", dG[e:e + 1])
    return (dx + dx0).astype(dout.dtype)


class DCN(Base):
    """model/dcn.py:12-43."""

    def __init__(self, field_dims, embed_dim, n_cross_layers, mlp_dims,
                 l2_reg_embedding=1e-5, l2_reg_linear=1e-5, l2_reg_dnn=1e-5, l2_reg_cross=1e-5, **_):
        super().__init__(field_dims, embed_dim, l2_reg_embedding, l2_reg_linear)
        self.L = n_cross_layers
        self.mlp = MLP("mlp", len(mlp_dims), True, False)
        self.reg_prefixes = (("mlp", l2_reg_dnn), ("cn", l2_reg_cross))

    def forward(self, sd, x, train=True):
        bufs = {}
        e, idx = self.embed(sd, x)
        cn, xs = cross_v1_fwd(sd, "cn", self.L, e)
        mo, mc = self.mlp.forward(sd, e, train, bufs, self.drop)
        st = np.concatenate([cn, mo], axis=1)
        y = sigmoid(self.lin_fwd(sd, e) + st @ sd["mlp_linear.weight"].T)
        return y[:, 0], dict(idx=idx, e=e, xs=xs, mc=mc, st=st, y=y, bufs=bufs)

    def backward(self, sd, cache, dy):
        grads = {}
        y, e = cache["y"], cache["e"]
        dz = (dy.reshape(-1, 1) * y * (1 - y)).astype(e.dtype)
        dst, dW, _ = linear_bwd(cache["st"], sd["mlp_linear.weight"], dz, has_bias=False)
        _acc(grads, "mlp_linear.weight", dW)
        D = self.D
        dembed = cross_v1_bwd(sd, "cn", self.L, cache["xs"], dst[:, :D], grads)
        dembed = dembed + self.mlp.backward(sd, cache["mc"], dst[:, D:], grads)
        self.embed_bwd(sd, cache["idx"], e, dembed, dz, grads)
        return grads


class AutoInt(Base):
    """model/autoint.py:10-64: the field self-attention block (same stages as BaseModel.atten_forward, layer.py:71-84) whose
    relu(output).view(B, F*A) is concatenated with an MLP over the flat embeddings; one bias-free Linear over the concatenation,
    plus FeaturesLinear, sigmoid."""

    def __init__(self, field_dims, embed_dim, atten_embed_dim=None, att_layer_num=3, att_head_num=2, att_res=True, mlp_dims=(256, 128),
                 l2_reg_embedding=1e-5, l2_reg_linear=1e-5, l2_reg_dnn=1e-5, **_):
        super().__init__(field_dims, embed_dim, l2_reg_embedding, l2_reg_linear)
        A = embed_dim if atten_embed_dim is None else atten_embed_dim
        self.enable_atten(A, att_layer_num, att_head_num, att_res, head=("dnn_linear.weight", 0))
        self.FA = self.F * A
        self.dnn = MLP("dnn", len(mlp_dims), True, False)
        self.reg_prefixes = (("dnn", l2_reg_dnn),)

    def forward(self, sd, x, train=True):
        bufs = {}
        e, idx = self.embed(sd, x)
        att, ac = self.atten_fwd(sd, e)                                   # relu(cross_term) . dnn_linear.weight[:, :F*A]
        mo, mc = self.dnn.forward(sd, e, train, bufs, self.drop)
        lin = linear_fwd(e, sd["linear.fc.weight"], sd["linear.fc.bias"])
        y = sigmoid(lin + att + mo @ sd["dnn_linear.weight"][:, self.FA:].T)
        return y[:, 0], dict(idx=idx, e=e, ac=ac, mc=mc, mo=mo, y=y, bufs=bufs)

    def backward(self, sd, cache, dy):
        grads = {}
        y, e = cache["y"], cache["e"]
        dz = (dy.reshape(-1, 1) * y * (1 - y)).astype(e.dtype)
        dmo, dW, _ = linear_bwd(cache["mo"], sd["dnn_linear.weight"][:, self.FA:], dz, has_bias=False)
        dfull = np.zeros_like(sd["dnn_linear.weight"])
        dfull[:, self.FA:] = dW
        _acc(grads, "dnn_linear.weight", dfull)
        dembed = self.dnn.backward(sd, cache["mc"], dmo, grads)
        dembed = dembed + self.atten_bwd(sd, cache["ac"], dz, grads)
        self._att_cache = None
        self.embed_bwd(sd, cache["idx"], e, dembed, dz, grads)
        return grads


class DCNv2(Base):
    """model/dcnv2.py:9-70 (CrossNetMix default) plus the CrossNetV2 variant that upstream's
    constructor cannot build assembled from the reference layers."""

    def __init__(self, field_dims, embed_dim, n_cross_layers, mlp_dims, model_structure="parallel",
                 use_low_rank_mixture=True, num_experts=4,
                 l2_reg_embedding=1e-5, l2_reg_linear=1e-5, l2_reg_dnn=1e-5, l2_reg_cross=1e-5, **_):
        super().__init__(field_dims, embed_dim, l2_reg_embedding, l2_reg_linear)
        self.L, self.struct, self.mix, self.ne = n_cross_layers, model_structure, use_low_rank_mixture, num_experts
        self.dnn = MLP("dnn", len(mlp_dims), True, False)
        self.reg_prefixes = (("dnn", l2_reg_dnn),)
        ex = [("dnn_linear.weight", l2_reg_linear)]
        if use_low_rank_mixture:
            for l in range(n_cross_layers):
                ex += [(f"crossnet.u_list.{l}", l2_reg_cross), (f"crossnet.v_list.{l}", l2_reg_cross),
                       (f"crossnet.c_list.{l}", l2_reg_cross)]
        self.reg_exact = tuple(sorted(ex, key=lambda kv: (0 if kv[0].startswith("dnn_linear") else 1,
                                                             ["u", "v", "c", "d"].index(kv[0].split(".")[1][0]) if "crossnet" in kv[0] else 0)))

    def forward(self, sd, x, train=True):
        bufs = {}
        e, idx = self.embed(sd, x)
        if self.mix:
            co, cc = cross_mix_fwd(sd, "crossnet", self.L, self.ne, e)
        else:
            co, cc = cross_v2_fwd(sd, "crossnet", self.L, e)
        if self.struct == "parallel":
            do, dc = self.dnn.forward(sd, e, train, bufs, self.drop)
            fin = np.concatenate([co, do], axis=1)
        else:                                              # stacked
            do, dc = self.dnn.forward(sd, co, train, bufs, self.drop)
            fin = do
        y = sigmoid(fin @ sd["dnn_linear.weight"].T + self.lin_fwd(sd, e))
        return y[:, 0], dict(idx=idx, e=e, cc=cc, dc=dc, fin=fin, y=y, bufs=bufs)

    def _cross_bwd(self, sd, cache, e, d, grads):
        if self.mix:
            return cross_mix_bwd(sd, "crossnet", self.L, self.ne, e, cache["cc"], d, grads)
        return cross_v2_bwd(sd, "crossnet", self.L, cache["cc"], d, grads)

    def backward(self, sd, cache, dy):
        grads = {}
        y, e = cache["y"], cache["e"]
        dz = (dy.reshape(-1, 1) * y * (1 - y)).astype(e.dtype)
        dfin, dW, _ = linear_bwd(cache["fin"], sd["dnn_linear.weight"], dz, has_bias=False)
        _acc(grads, "dnn_linear.weight", dW)
        if self.struct == "parallel":
            dembed = self._cross_bwd(sd, cache, e, dfin[:, :self.D], grads)
            dembed = dembed + self.dnn.backward(sd, cache["dc"], dfin[:, self.D:], grads)
        else:
            dco = self.dnn.backward(sd, cache["dc"], dfin, grads)
            dembed = self._cross_bwd(sd, cache, e, dco, grads)
        self.embed_bwd(sd, cache["idx"], e, dembed, dz, grads)
        return grads


# --------------------------------------------------------------------------------------
# STAR (model/star.py:12-187)
# --------------------------------------------------------------------------------------
def route_partition(group, n_group):
    """star.py:84-86,112-114: rows of group g keep their original order; groups are concatenated
    in ascending g.  Returns (perm, counts): perm[i] = source row of output row i."""
    group = np.asarray(group).reshape(-1)
    perm = np.concatenate([np.flatnonzero(group == g) for g in range(n_group)]) if n_group else np.zeros(0, np.int64)
    counts = np.array([(group == g).sum() for g in range(n_group)], dtype=np.int64)
    return perm.astype(np.int64), counts


class STAR(Base):
    def __init__(self, field_dims, embed_dim, n_tower, tower_dims,
                 l2_reg_embedding=1e-5, l2_reg_linear=1e-5, l2_reg_dnn=1e-5, **_):
        super().__init__(field_dims, embed_dim, l2_reg_embedding, l2_reg_linear)
        self.T, self.nl = n_tower, len(tower_dims)
        self.reg_prefixes = (("domain_dnns", l2_reg_dnn), ("shared_dnn", l2_reg_dnn))

    def _tower_fwd(self, sd, g, xin, train, bufs):
        """star.py:78-103 for one tower on its rows."""
        cache = {}
        n = xin.shape[0]
        k = f"domain_norm.{g}"
        if n == 1:                                       # star.py:134-135
            h = xin; cache["pn"] = None
        else:
            gam = sd[k + ".weight"] * sd["shared_bn_weight"]
            bet = sd[k + ".bias"] + sd["shared_bn_bias"]
            h, bc, upd = bn_fwd(xin, gam, bet, sd[k + ".running_mean"], sd[k + ".running_var"], train)
            cache["pn"] = bc
            if upd is not None:
                bufs[k + ".running_mean"], bufs[k + ".running_var"] = upd
        if train:                                        # star.py:146-148 runs before the size test? no: after the B==1 return
            if n != 1:
                bufs[k + ".num_batches_tracked"] = sd[k + ".num_batches_tracked"] + 1
        layers = []
        for i in range(self.nl):
            Wd, Ws = sd[f"domain_dnns.{g}.linears.{i}.weight"], sd[f"shared_dnn.linears.{i}.weight"]
            bd, bs = sd[f"domain_dnns.{g}.linears.{i}.bias"], sd[f"shared_dnn.linears.{i}.bias"]
            z = linear_fwd(h, Wd * Ws, bd + bs)
            bc = None
            if z.shape[0] > 1:                           # star.py:94
                kb = f"domain_dnns.{g}.bn.{i}"
                z, bc, upd = bn_fwd(z, sd[kb + ".weight"], sd[kb + ".bias"], sd[kb + ".running_mean"],
                                    sd[kb + ".running_var"], train)
                if upd is not None:
                    bufs[kb + ".running_mean"], bufs[kb + ".running_var"] = upd
                    bufs[kb + ".num_batches_tracked"] = sd[kb + ".num_batches_tracked"] + 1
            a = np.maximum(z, 0)
            layers.append((h, bc, a > 0))
            h = a
        Wd, Ws = sd[f"domain_dnn_linears.{g}.weight"], sd["shared_dnn_linear.weight"]
        logit = linear_fwd(h, Wd * Ws, sd[f"domain_dnn_linears.{g}.bias"] + sd["shared_dnn_linear.bias"])
        cache.update(xin=xin, layers=layers, hlast=h, train=train)
        return logit, cache

    def _tower_bwd(self, sd, g, cache, dlogit, grads):
        train = cache["train"]
        Wd, Ws = sd[f"domain_dnn_linears.{g}.weight"], sd["shared_dnn_linear.weight"]
        dh, dW, db = linear_bwd(cache["hlast"], Wd * Ws, dlogit)
        _acc(grads, f"domain_dnn_linears.{g}.weight", dW * Ws); _acc(grads, "shared_dnn_linear.weight", dW * Wd)
        _acc(grads, f"domain_dnn_linears.{g}.bias", db); _acc(grads, "shared_dnn_linear.bias", db)
        for i in reversed(range(self.nl)):
            h, bc, mask = cache["layers"][i]
            dz = dh * mask
            if bc is not None:
                kb = f"domain_dnns.{g}.bn.{i}"
                dz, dg, dbt = bn_bwd(dz, bc, train)
                _acc(grads, kb + ".weight", dg); _acc(grads, kb + ".bias", dbt)
            else:                                          # keep zero-grad semantics only when BN ran
                pass
            Wd, Ws = sd[f"domain_dnns.{g}.linears.{i}.weight"], sd[f"shared_dnn.linears.{i}.weight"]
            dh, dW, db = linear_bwd(h, Wd * Ws, dz)
            _acc(grads, f"domain_dnns.{g}.linears.{i}.weight", dW * Ws); _acc(grads, f"shared_dnn.linears.{i}.weight", dW * Wd)
            _acc(grads, f"domain_dnns.{g}.linears.{i}.bias", db); _acc(grads, f"shared_dnn.linears.{i}.bias", db)
        if cache["pn"] is not None:
            k = f"domain_norm.{g}"
            dx, dgam, dbet = bn_bwd(dh, cache["pn"], train)
            _acc(grads, k + ".weight", dgam * sd["shared_bn_weight"]); _acc(grads, "shared_bn_weight", dgam * sd[k + ".weight"])
            _acc(grads, k + ".bias", dbet); _acc(grads, "shared_bn_bias", dbet)
            return dx
        return dh

    def forward(self, sd, x, train=True, group=None):
        """star.py:62-114.  group=None: every tower sees every row -> (B,T).  With group: rows are
        partitioned (stable) and the result is (B,1) in partition order (+ the permutation)."""
        bufs = {}
        e, idx = self.embed(sd, x)
        lin = self.lin_fwd(sd, e)
        ys, tcs = [], []
        if group is None:
            for g in range(self.T):
                lg, c = self._tower_fwd(sd, g, e, train, bufs)
                ys.append(sigmoid(lg + lin)); tcs.append(c)
            y = np.concatenate(ys, axis=1)
            return y, dict(idx=idx, e=e, tcs=tcs, y=y, bufs=bufs, perm=None)
        perm, counts = route_partition(group, self.T)
        off = 0
        for g in range(self.T):
            rows = perm[off:off + counts[g]]; off += counts[g]
            lg, c = self._tower_fwd(sd, g, e[rows], train, bufs)
            ys.append(sigmoid(lg + lin[rows])); tcs.append(c)
        y = np.concatenate(ys, axis=0)
        return y, dict(idx=idx, e=e, tcs=tcs, y=y, bufs=bufs, perm=perm, counts=counts)

    def backward(self, sd, cache, dy):
        grads = {}
        e, y = cache["e"], cache["y"]
        dlogit = (dy.reshape(y.shape) * y * (1 - y)).astype(e.dtype)
        dembed = np.zeros_like(e)
        if cache["perm"] is None:
            for g in range(self.T):
                dembed += self._tower_bwd(sd, g, cache["tcs"][g], dlogit[:, g:g + 1], grads)
            dlin = dlogit.sum(1, keepdims=True)
        else:
            perm, counts = cache["perm"], cache["counts"]
            dlin = np.zeros((e.shape[0], 1), dtype=e.dtype)
            off = 0
            for g in range(self.T):
                rows = perm[off:off + counts[g]]
                dl = dlogit[off:off + counts[g]]; off += counts[g]
                dembed[rows] += self._tower_bwd(sd, g, cache["tcs"][g], dl, grads)
                dlin[rows] += dl
        self.embed_bwd(sd, cache["idx"], e, dembed, dlin, grads)
        return grads


# --------------------------------------------------------------------------------------
# CDC tower selection (model/cdc.py:95-111) and the train step (run.py:483-492, 635-640)
# --------------------------------------------------------------------------------------
def select_pred(y, mode, group=None, domain2group=None, x=None, domain_idx=None, domain_i=None):
    """Returns (selected prediction (B,), function mapping d(selected) -> d(y))."""
    B = y.shape[0]
    if mode == "single":                                  # run.py:486-488
        return y.reshape(B), lambda d: d.reshape(y.shape)
    if mode == "gather":                                  # run.py:483-484
        g = np.asarray(group).reshape(B)
    elif mode == "warmup":                                # cdc.py:99-102
        T = y.shape[1]
        return y.mean(1), lambda d: np.repeat(d.reshape(B, 1) / T, T, axis=1).astype(y.dtype)
    elif mode == "split_domain":                          # cdc.py:108-111
        g = np.full(B, int(np.asarray(domain2group)[domain_i]))
    elif mode == "split_gather":                          # cdc.py:104-107
        g = np.asarray(domain2group)[x[:, domain_idx].astype(np.int64)]
    else:
        raise ValueError(mode)

    def back(d):
        dy = np.zeros_like(y)
        dy[np.arange(B), g] = d
        return dy
    return y[np.arange(B), g], back


class Adam:
    """torch.optim.Adam single-tensor path (run.py:720-721): L2-style weight decay
    added to the gradient, beta=(0.9,0.99), eps 1e-8; parameters whose grad is None are skipped
    (their per-parameter step count does not advance)."""

    def __init__(self, lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8):
        self.lr, self.b1, self.b2, self.eps, self.wd = lr, betas[0], betas[1], eps, weight_decay
        self.m, self.v, self.t = {}, {}, {}

    def step(self, sd, grads):
        for k, g in grads.items():
            p = sd[k]
            if k not in self.m:
                self.m[k] = np.zeros_like(p); self.v[k] = np.zeros_like(p); self.t[k] = 0
            self.t[k] += 1
            t = self.t[k]
            g = g.astype(p.dtype).reshape(p.shape) + F32(self.wd) * p
            m, v = self.m[k], self.v[k]
            m += F32(1 - self.b1) * (g - m)
            v *= F32(self.b2); v += F32(1 - self.b2) * g * g
            bc1 = 1 - self.b1 ** t
            bc2s = np.sqrt(1 - self.b2 ** t)
            denom = np.sqrt(v) / F32(bc2s) + F32(self.eps)
            sd[k] = (p - F32(self.lr / bc1) * (m / denom)).astype(p.dtype)


def train_step(model, sd, opt, x, y, mode, **sel):
    """One pass of run.py:483-492 (or 635-640 for CDC).  Mutates sd (params + BN buffers).
    Returns dict(pred, bce, reg, loss, grads)."""
    fkw = {"group": sel["group"]} if (isinstance(model, STAR) and mode == "star_grouped") else {}
    pred, cache = model.forward(sd, x, train=True, **fkw)
    if mode == "star_grouped":
        tgt = np.asarray(y).reshape(-1)[cache["perm"]].astype(np.float64)
        psel, back = pred.reshape(-1), (lambda d: d.reshape(pred.shape))
    else:
        tgt = np.asarray(y).reshape(-1).astype(np.float64)
        psel, back = select_pred(pred, mode, x=x, **sel)
    bce, dp = bce_mean(psel, tgt)
    reg = model.reg_loss(sd)
    grads = model.backward(sd, cache, back(dp))
    model.add_reg_grad(sd, grads)
    for k, v in cache["bufs"].items():
        sd[k] = v
    if opt is not None:
        opt.step(sd, grads)
    return dict(pred=pred, psel=psel, bce=bce, reg=reg, loss=F32(bce + reg), grads=grads)


# --------------------------------------------------------------------------------------
# N4: distance-covariance causal kernel (model/cdc.py:364-393), host-side float64
# --------------------------------------------------------------------------------------
def calc_causal_matrix(X):
    """cdc.py:364-393 with alpha=None: Z_j doubly-centred |x_i-x_k| per feature j; gamma =
    (F^T F)^2 - 2 <Z,Z>_I + ||I||_F ; kappa = cosine-normalised gamma clipped at 1."""
    X = np.asarray(X, dtype=np.float64)
    n, f = X.shape
    Z = np.zeros((f, n, n))
    for j in range(f):
        D = np.abs(X[:, j][:, None] - X[:, j][None, :])
        Z[j] = (D - D.mean(0) - D.mean(1).reshape(-1, 1)) / D.mean() + 1
    Fm = Z.reshape(f * n, n)
    zz = np.einsum("jab,jbc->ac", Z, Z)                  # thresh = I: sum_j Z_j Z_j
    gamma = (Fm.T @ Fm) ** 2 - 2 * zz + np.sqrt(f)      # ||I_f||_F = sqrt(f)
    d = np.diag(gamma)
    kappa = gamma / np.sqrt(np.outer(d, d))
    kappa[kappa > 1] = 1
    return kappa


def strip_prefix(sd, prefix="base_model_instance."):
    return {k[len(prefix):] if k.startswith(prefix) else k: v for k, v in sd.items()}
