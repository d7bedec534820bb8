"""The field self-attention block of BaseModel (reference model/layer.py:58-84): with `config.use_atten`
(on in the stock config.py:24-28) PLE / MMoE / STAR add one more scalar to every tower logit,

    atten_x = atten_embedding(embed_x.view(B, F, E))                     Linear(E -> A) per field token
    for attn in self_attns:  x = MultiheadAttention(A, heads)(x, x, x)    in_proj -> softmax(q k^T / sqrt(dh)) v -> out_proj
    x = relu(x + V_res_embedding(embed_x))                               (att_res)
    out = atten_linear(x.view(B, F*A))                                   Linear(F*A -> 1, no bias)

Everything acts on the token matrix [B*F, .] (row b*F + f); the reference's transposes are views.  fp32 path: the Linear layers
run on the fp32 GEMM entry point, the attention core and the ReLU + Linear head are cdcmdr_attn_* (csrc/attn.cu).  bf16 path
(`can_bf16`): the gathered bf16 embeddings ARE the token matrix [B*F, E]; every projection, input gradient and weight gradient is
a tcgen05 GEMM over bf16 token matrices (cdcmdr_gemm_bf16_tc: fp32 accumulation, fp32 weight gradients by split-K over the B*F
rows), the core and the head are cdcmdr_attn_*_bf16 (fp32 arithmetic inside, no stored probabilities).  The scalar lands in the same buffer as the FeaturesLinear logit (both are
`other_outs` added to every tower, layer.py:52-54), so the head kernel accumulates onto it and the backward reads the same dlin."""
from __future__ import annotations

import math

from torch import nn

from .core import Mat


def build_atten(model, config, dropout):
    """layer.py:58-69 - the parameter holders, created in the reference's order with the reference's names / shapes / init."""
    A = getattr(config, 'atten_embed_dim', model.embed_dim)
    if config.atten_embed_dim != A:
        raise ValueError("atten_embed_dim mismatch")
    model.atten_embedding = nn.Linear(model.embed_dim, A)
    model.atten_output_dim = model.embedding.output_dim0 * A
    model.att_res = config.att_res
    model.self_attns = nn.ModuleList([nn.MultiheadAttention(config.atten_embed_dim, config.att_head_num, dropout=dropout)
                                      for _ in range(config.att_layer_num)])
    if model.att_res:
        model.V_res_embedding = nn.Linear(model.embed_dim, A)
    model.atten_linear = nn.Linear(model.atten_output_dim, 1, bias=False)
    if config.att_layer_num < 1:
        raise NotImplementedError("cdcmdr: the attention block needs att_layer_num >= 1")
    if model.field_num > 32:
        raise NotImplementedError("cdcmdr: the attention core handles at most 32 field tokens")
    model._att_geom = (A, int(config.att_head_num), int(config.att_layer_num))


class AttnBlock:
    def __init__(self, model, rt):
        self.rt = rt
        self.F, self.E = model.field_num, model.embed_dim
        self.A, self.H, self.n_layer = model._att_geom
        self.dh = self.A // self.H
        self.res = bool(model.att_res)
        self.scale = 1.0 / math.sqrt(self.dh)
        self.salts = [0xA77E0000 + i for i in range(self.n_layer)]
        # the bias-free Linear over relu(block output).view(B, F*A): `atten_linear` for BaseModel.atten_forward (layer.py:82-83), the
        # first F*A columns of `dnn_linear` for AutoInt (autoint.py:60-62)
        self.head_name, self.head_off = getattr(model, "_att_head", ("atten_linear.weight", 0))
        # tensor-core path: TMA needs 16-byte aligned bf16 pitches (E, A multiples of 8), the core loads bf16 pairs
        self.can_bf16 = bool(rt.bf16 and self.E % 8 == 0 and self.A % 8 == 0 and self.dh % 2 == 0)

    # ---------------------------------------------------------------- helpers
    def _tok(self, X: Mat, B) -> Mat:
        if X.ld != self.F * self.E or (X.is_bf16 and not self.can_bf16):
            raise RuntimeError("cdcmdr: the attention block reads the embeddings [B, F*E] as contiguous tokens")
        return Mat(X.t, X.off, self.E)

    def _drop(self, train):
        rt = self.rt
        p = rt.dropout if train else 0.0
        return p, (rt.seed_ptr if p > 0 else None)

    # ---------------------------------------------------------------- forward: lin[b] += atten_forward(embed_x)[b]
    def fwd(self, ws, X32: Mat, B, lin: Mat, train):
        if X32.is_bf16:
            return self._fwd_bf16(ws, X32, B, lin, train)
        rt, ops = self.rt, self.rt.ops
        F, E, A, H, dh = self.F, self.E, self.A, self.H, self.dh
        M = B * F
        Xt = self._tok(X32, B)
        cur = ws.mat("att.t0", M, A)
        rt.lin_fwd(Xt, E, rt.o("atten_embedding.weight"), A, rt.o("atten_embedding.bias"), cur, M)
        p, seed = self._drop(train)
        for i in range(self.n_layer):
            pre = f"self_attns.{i}."
            qkv = ws.mat(f"att.qkv{i}", M, 3 * A)
            rt.lin_fwd(cur, A, rt.o(pre + "in_proj_weight"), 3 * A, rt.o(pre + "in_proj_bias"), qkv, M)
            o = ws.mat(f"att.o{i}", M, A)
            probs = ws.get(f"att.p{i}", (B * H * F * F,))
            ops.lib.attn_fwd(qkv.ptr, qkv.ld, o.ptr, o.ld, probs.data_ptr(), B, F, H, dh, self.scale, p, seed, self.salts[i], ops.stream)
            y = ws.mat(f"att.y{i}", M, A)
            if i + 1 == self.n_layer and self.res:
                rt.lin_fwd(Xt, E, rt.o("V_res_embedding.weight"), A, rt.o("V_res_embedding.bias"), y, M)
                ops.gemm_f32(A=o.ptr, a_rs=o.ld, a_cs=1, Bt=rt.w(pre + "out_proj.weight"), b_rs=A, b_cs=1, Cm=y.ptr, c_rs=y.ld,
                             M=M, N=A, K=A, bias=rt.w(pre + "out_proj.bias"), accumulate=1)
            else:
                rt.lin_fwd(o, A, rt.o(pre + "out_proj.weight"), A, rt.o(pre + "out_proj.bias"), y, M)
            cur = y
        ops.lib.attn_pool_fwd(cur.ptr, rt.w(self.head_name, self.head_off), lin.ptr, lin.ld, 1, B, F * A, ops.stream)

    # ---------------------------------------------------------------- backward: parameter gradients; dX += d(embed_x)
    def bwd(self, ws, X32: Mat, B, dlin: Mat, dX: Mat, train):
        if X32.is_bf16:
            return self._bwd_bf16(ws, X32, B, dlin, dX, train)
        rt, ops = self.rt, self.rt.ops
        F, E, A, H, dh = self.F, self.E, self.A, self.H, self.dh
        M = B * F
        Xt = self._tok(X32, B)
        if dX.is_bf16 or dX.ld != F * E:
            raise RuntimeError("cdcmdr: the attention block adds into the fp32 embedding gradient [B, F*E]")
        dXt = Mat(dX.t, dX.off, E)
        p, seed = self._drop(train)
        z = ws.mat(f"att.y{self.n_layer - 1}", M, A)
        dy = ws.mat("att.dz", M, A)
        sc = ops.scratch("attn_pool", ops.lib.attn_pool_scratch_bytes(B, F * A))
        ops.lib.attn_pool_bwd(z.ptr, rt.w(self.head_name, self.head_off), dlin.ptr, dlin.ld, dy.ptr, rt.g(self.head_name, self.head_off), B, F * A,
                              sc.data_ptr(), ops.stream)
        if self.res:
            ops.colsum(dy, M, A, rt.g("V_res_embedding.bias"))
            rt.lin_bwd_w(dy, Xt, E, rt.o("V_res_embedding.weight"), A, M)
            rt.lin_bwd_x(dy, E, rt.o("V_res_embedding.weight"), A, dXt, M, accumulate=True)
        for i in reversed(range(self.n_layer)):
            pre = f"self_attns.{i}."
            qkv, o = ws.mat(f"att.qkv{i}", M, 3 * A), ws.mat(f"att.o{i}", M, A)
            probs = ws.get(f"att.p{i}", (B * H * F * F,))
            src = ws.mat("att.t0", M, A) if i == 0 else ws.mat(f"att.y{i - 1}", M, A)
            ops.colsum(dy, M, A, rt.g(pre + "out_proj.bias"))
            rt.lin_bwd_w(dy, o, A, rt.o(pre + "out_proj.weight"), A, M)
            do = ws.mat("att.do", M, A)
            rt.lin_bwd_x(dy, A, rt.o(pre + "out_proj.weight"), A, do, M)
            dqkv = ws.mat("att.dqkv", M, 3 * A)
            ops.lib.attn_bwd(qkv.ptr, qkv.ld, probs.data_ptr(), do.ptr, do.ld, dqkv.ptr, dqkv.ld, B, F, H, dh, self.scale, p, seed,
                             self.salts[i], ops.stream)
            ops.colsum(dqkv, M, 3 * A, rt.g(pre + "in_proj_bias"))
            rt.lin_bwd_w(dqkv, src, A, rt.o(pre + "in_proj_weight"), 3 * A, M)
            dy = ws.mat("att.dcur", M, A)
            rt.lin_bwd_x(dqkv, A, rt.o(pre + "in_proj_weight"), 3 * A, dy, M)
        ops.colsum(dy, M, A, rt.g("atten_embedding.bias"))
        rt.lin_bwd_w(dy, Xt, E, rt.o("atten_embedding.weight"), A, M)
        rt.lin_bwd_x(dy, E, rt.o("atten_embedding.weight"), A, dXt, M, accumulate=True)

    # ---------------------------------------------------------------- the same block on bf16 token matrices (tcgen05 GEMMs)
    def _fwd_bf16(self, ws, X: Mat, B, lin: Mat, train):
        import torch
        rt, ops = self.rt, self.rt.ops
        F, E, A, H, dh = self.F, self.E, self.A, self.H, self.dh
        M, bf = B * F, torch.bfloat16
        Xt = self._tok(X, B)
        cur = ws.mat("att.t0", M, A, bf)
        rt.lin_fwd(Xt, E, rt.o("atten_embedding.weight"), A, rt.o("atten_embedding.bias"), cur, M)
        p, seed = self._drop(train)
        for i in range(self.n_layer):
            pre = f"self_attns.{i}."
            qkv = ws.mat(f"att.qkv{i}", M, 3 * A, bf)
            rt.lin_fwd(cur, A, rt.o(pre + "in_proj_weight"), 3 * A, rt.o(pre + "in_proj_bias"), qkv, M)
            o = ws.mat(f"att.o{i}", M, A, bf)
            ops.lib.attn_fwd_bf16(qkv.ptr, qkv.ld, o.ptr, o.ld, B, F, H, dh, self.scale, p, seed, self.salts[i], ops.stream)
            y = ws.mat(f"att.y{i}", M, A, bf)
            if i + 1 == self.n_layer and self.res:                 # y = V_res(x) + out_proj(o): the second GEMM accumulates onto the first
                rt.lin_fwd(Xt, E, rt.o("V_res_embedding.weight"), A, rt.o("V_res_embedding.bias"), y, M)
                ops.gemm_tc(A=o.ptr, lda=o.ld, a_rows=M, a_cols=A, a_mn=0, Bt=rt.Wb.data_ptr() + 2 * rt.o(pre + "out_proj.weight"), ldb=A,
                            b_rows=A, b_cols=A, b_mn=0, M=M, N=A, K=A, bias=rt.w(pre + "out_proj.bias"), n_main=A, out_main=y.ptr,
                            ld_main=y.ld, accumulate=1)
            else:
                rt.lin_fwd(o, A, rt.o(pre + "out_proj.weight"), A, rt.o(pre + "out_proj.bias"), y, M)
            cur = y
        ops.lib.attn_pool_fwd_bf16(cur.ptr, rt.w(self.head_name, self.head_off), lin.ptr, lin.ld, 1, B, F * A, ops.stream)

    def _bwd_bf16(self, ws, X: Mat, B, dlin: Mat, dX: Mat, train):
        import torch
        rt, ops = self.rt, self.rt.ops
        F, E, A, H, dh = self.F, self.E, self.A, self.H, self.dh
        M, bf = B * F, torch.bfloat16
        Xt = self._tok(X, B)
        if dX.is_bf16 or dX.ld != F * E:
            raise RuntimeError("cdcmdr: the attention block adds into the fp32 embedding gradient [B, F*E]")
        dXt = Mat(dX.t, dX.off, E)
        p, seed = self._drop(train)
        z = ws.mat(f"att.y{self.n_layer - 1}", M, A, bf)
        dy = ws.mat("att.dz", M, A, bf)
        sc = ops.scratch("attn_pool", ops.lib.attn_pool_scratch_bytes(B, F * A))
        ops.lib.attn_pool_bwd_bf16(z.ptr, rt.w(self.head_name, self.head_off), dlin.ptr, dlin.ld, dy.ptr, rt.g(self.head_name, self.head_off), B, F * A,
                                   sc.data_ptr(), ops.stream)
        if self.res:
            ops.colsum(dy, M, A, rt.g("V_res_embedding.bias"))
            rt.lin_bwd_w(dy, Xt, E, rt.o("V_res_embedding.weight"), A, M)
            rt.lin_bwd_x(dy, E, rt.o("V_res_embedding.weight"), A, dXt, M, accumulate=True)
        for i in reversed(range(self.n_layer)):
            pre = f"self_attns.{i}."
            qkv, o = ws.mat(f"att.qkv{i}", M, 3 * A, bf), ws.mat(f"att.o{i}", M, A, bf)
            src = ws.mat("att.t0", M, A, bf) if i == 0 else ws.mat(f"att.y{i - 1}", M, A, bf)
            ops.colsum(dy, M, A, rt.g(pre + "out_proj.bias"))
            rt.lin_bwd_w(dy, o, A, rt.o(pre + "out_proj.weight"), A, M)
            do = ws.mat("att.do", M, A, bf)
            rt.lin_bwd_x(dy, A, rt.o(pre + "out_proj.weight"), A, do, M)
            dqkv = ws.mat("att.dqkv", M, 3 * A, bf)
            ops.lib.attn_bwd_bf16(qkv.ptr, qkv.ld, do.ptr, do.ld, dqkv.ptr, dqkv.ld, B, F, H, dh, self.scale, p, seed, self.salts[i],
                                  ops.stream)
            ops.colsum(dqkv, M, 3 * A, rt.g(pre + "in_proj_bias"))
            rt.lin_bwd_w(dqkv, src, A, rt.o(pre + "in_proj_weight"), 3 * A, M)
            dy = ws.mat("att.dcur", M, A, bf)
            rt.lin_bwd_x(dqkv, A, rt.o(pre + "in_proj_weight"), 3 * A, dy, M)
        ops.colsum(dy, M, A, rt.g("atten_embedding.bias"))
        rt.lin_bwd_w(dy, Xt, E, rt.o("atten_embedding.weight"), A, M)
        rt.lin_bwd_x(dy, E, rt.o("atten_embedding.weight"), A, dXt, M, accumulate=True)
