"""DCN (reference model/dcn.py:12-43) and DCNv2 (model/dcnv2.py:9-70) on libcdcmdr.so.

  DCN    sigmoid( linear(e) + [CrossNetwork(e) | MLP(e)] . w )                          cross: layer.py:495-515 (live def.)
  DCNv2  sigmoid( linear(e) + [CrossNetMix(e) | MLP(e)] . w )   (parallel)              cross: layer.py:346-407
         sigmoid( linear(e) + MLP(CrossNetMix(e)) . w )          (stacked)
         use_low_rank_mixture=False selects CrossNetV2 (layer.py:332-343; x0*(W x) + b + x, bias OUTSIDE the product,
         SURVEY G8).  Upstream's constructor crashes for that flag; here it builds the layer the flag names.

The MLP branch is an `MlpGroup` (tensor-core GEMMs on the bf16 path); the cross networks run in fp32: their GEMM-shaped
pieces through cdcmdr_gemm_f32 with the operands read in place (strides), the Hadamard / gating stages through the fused
elementwise kernels of ops_cross_route.cu.  Single-logit models: forward returns (B,) probabilities; the fused step is
`train_step(x, y, optimizer, mode="col", col=0)`."""
from __future__ import annotations

import torch
from torch import nn

from .core import Mat
from .layer import BaseModel, MultiLayerPerceptron, mlp_group_names, precision_of
from .runtime import MlpGroup


# --------------------------------------------------------------------------------------------- parameter holders
class CrossNetwork(nn.Module):
    """layer.py:495-515: w.N = Linear(D, 1, bias=False), b.N = (D,) zeros."""

    def __init__(self, input_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        self.w = nn.ModuleList([nn.Linear(input_dim, 1, bias=False) for _ in range(num_layers)])
        self.b = nn.ParameterList([nn.Parameter(torch.zeros((input_dim,))) for _ in range(num_layers)])


class CrossNetV2(nn.Module):
    """layer.py:332-343: w.N = Linear(D, D, bias=False), b.N = (D,) zeros."""

    def __init__(self, input_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        self.w = nn.ModuleList([nn.Linear(input_dim, input_dim, bias=False) for _ in range(num_layers)])
        self.b = nn.ParameterList([nn.Parameter(torch.zeros((input_dim,))) for _ in range(num_layers)])


class CrossNetMix(nn.Module):
    """layer.py:346-407: per layer U, V (n_exp, D, r), C (n_exp, r, r), bias (D, 1); gating.e = Linear(D, 1, bias=False)
    shared by all layers."""

    def __init__(self, input_dim, num_layers=2, low_rank=32, num_experts=4):
        super().__init__()
        self.num_layers, self.num_experts, self.low_rank = num_layers, num_experts, low_rank
        self.u_list = nn.ParameterList([nn.Parameter(nn.init.xavier_normal_(torch.empty(num_experts, input_dim, low_rank)))
                                        for _ in range(num_layers)])
        self.v_list = nn.ParameterList([nn.Parameter(nn.init.xavier_normal_(torch.empty(num_experts, input_dim, low_rank)))
                                        for _ in range(num_layers)])
        self.c_list = nn.ParameterList([nn.Parameter(nn.init.xavier_normal_(torch.empty(num_experts, low_rank, low_rank)))
                                        for _ in range(num_layers)])
        self.gating = nn.ModuleList([nn.Linear(input_dim, 1, bias=False) for _ in range(num_experts)])
        self.bias = nn.ParameterList([nn.Parameter(nn.init.zeros_(torch.empty(input_dim, 1))) for _ in range(num_layers)])


# --------------------------------------------------------------------------------------------- shared program pieces
class _CrossModel(BaseModel):
    """Common tail of DCN / DCNv2: logit = linear(e) + cross_out . w[:D] + mlp_out . w[D:] (or mlp_out . w when stacked)."""

    n_out = 1

    def _shape_pred(self, pred, **kw):
        return pred[:, 0]                                           # y.squeeze(1)   dcn.py:43, dcnv2.py:70

    def _dlin_mat(self, ws, B):
        return ws.mat("dlin", B, 1)

    def _x32(self, ws, X: Mat, B) -> Mat:
        """fp32 view of the gathered embeddings for the cross network (the bf16 path gathers bf16 for the MLP GEMMs)."""
        if not X.is_bf16:
            return X
        D = self.embed_output_dim
        x32 = ws.mat("X32", B, D)
        self._rt.ops.cast_bf16_f32(X, x32, B, D)
        return x32

    # ---- wide linear  (layer.py:115-126)
    def _lin_fwd(self, ws, X: Mat, B) -> Mat:
        rt = self._rt
        lin = ws.mat("lin", B, 1)
        rt.ops.rowdot_fwd(X, rt.w("linear.fc.weight"), rt.w("linear.fc.bias"), lin, B, 1, self.embed_output_dim)
        return lin

    def _lin_bwd(self, ws, X: Mat, B, dX: Mat):
        rt = self._rt
        D = self.embed_output_dim
        tmp = ws.mat("dX.lin", B, D)
        rt.ops.rowdot_bwd(X, rt.w("linear.fc.weight"), self._dlin_mat(ws, B), tmp, rt.g("linear.fc.weight"), rt.g("linear.fc.bias"),
                          B, 1, D)
        rt.ops.add2d(tmp, dX, B, D, True)

    # ---- final Linear(D + h, 1, bias=False) over [cross | mlp]
    def _head_fwd(self, ws, wname, cross: Mat | None, mlp_out: Mat, B) -> Mat:
        rt = self._rt
        D, h = self.embed_output_dim, self._mlp.dims[-1]
        logit = ws.mat("head.logit", B, 1)
        if cross is None:
            rt.ops.rowdot_fwd(mlp_out, rt.w(wname), None, logit, B, 1, h)
            return logit
        rt.ops.rowdot_fwd(cross, rt.w(wname), None, logit, B, 1, D)
        part = ws.mat("head.logit_mlp", B, 1)
        rt.ops.rowdot_fwd(mlp_out, rt.w(wname, D), None, part, B, 1, h)
        rt.ops.add2d(part, logit, B, 1, True)
        return logit

    def _head_bwd(self, ws, wname, cross: Mat | None, mlp_out: Mat, dlogit: Mat, B):
        """-> (dcross fp32 [B, D] or None, dmlp fp32 [B, h])"""
        rt = self._rt
        D, h = self.embed_output_dim, self._mlp.dims[-1]
        dmlp = ws.mat("head.dmlp", B, h)
        if cross is None:
            rt.ops.rowdot_bwd(mlp_out, rt.w(wname), dlogit, dmlp, rt.g(wname), None, B, 1, h)
            return None, dmlp
        dcross = ws.mat("head.dcross", B, D)
        rt.ops.rowdot_bwd(cross, rt.w(wname), dlogit, dcross, rt.g(wname), None, B, 1, D)
        rt.ops.rowdot_bwd(mlp_out, rt.w(wname, D), dlogit, dmlp, rt.g(wname, D), None, B, 1, h)
        return dcross, dmlp


# --------------------------------------------------------------------------------------------- DCN
class DCN(_CrossModel):
    def __init__(self, feature_dims, embed_dim, n_cross_layers, mlp_dims, dropout=0.2,
                 l2_reg_embedding=1e-5, l2_reg_linear=1e-5, l2_reg_dnn=1e-5, l2_reg_cross=1e-5, config=None):
        super().__init__(feature_dims, embed_dim, l2_reg_embedding=l2_reg_embedding, l2_reg_linear=l2_reg_linear)
        self.model_name = 'dcn'
        self.n_cross_layers = n_cross_layers
        self.mlp_dims = tuple(mlp_dims)
        self.cn = CrossNetwork(self.embed_output_dim, n_cross_layers)
        self.mlp = MultiLayerPerceptron(self.embed_output_dim, mlp_dims, dropout, output_layer=False)
        self.mlp_linear = nn.Linear(self.embed_output_dim + mlp_dims[-1], 1, bias=False)
        self.output_layer = nn.Sigmoid()
        self.add_regularization_weight(self.reg_filter("mlp"), l2=l2_reg_dnn)
        self.add_regularization_weight(self.reg_filter("cn"), l2=l2_reg_cross)
        self._mlp_names, blk, bufs = mlp_group_names(["mlp"], self.mlp, "mlp")
        self._finalize(blk, bufs, precision=precision_of(config), dropout=dropout)

    def _on_runtime_built(self):
        self._mlp = MlpGroup(self._rt, "mlp", 1, self.embed_output_dim, self.mlp_dims, self._mlp_names, bn=True, out_layer=False,
                             in_groups=None)

    def _program_fwd(self, ws, X: Mat, B, train):
        rt, D, L = self._rt, self.embed_output_dim, self.n_cross_layers
        x0 = self._x32(ws, X, B)
        cur = x0
        for l in range(L):                                         # x <- x0 * (x . w_l) + b_l + x
            xw = ws.mat(f"cn.xw{l}", B, 1)
            rt.ops.rowdot_fwd(cur, rt.w(f"cn.w.{l}.weight"), None, xw, B, 1, D)
            nxt = ws.mat(f"cn.x{l + 1}", B, D)
            rt.ops.cross_fuse_fwd(x0, cur, xw, 1, rt.w(f"cn.b.{l}"), nxt, B, D)
            cur = nxt
        mlp_out = self._mlp.fwd(ws, X, B, train)
        logit = self._head_fwd(ws, "mlp_linear.weight", cur, mlp_out, B)
        return logit, self._lin_fwd(ws, X, B)

    def _program_bwd(self, ws, X: Mat, B, train, dlogits: Mat):
        rt, D, L = self._rt, self.embed_output_dim, self.n_cross_layers
        x0 = self._x32(ws, X, B) if not X.is_bf16 else ws.mat("X32", B, D)
        xs = [x0] + [ws.mat(f"cn.x{l + 1}", B, D) for l in range(L)]
        mlp_out = self._mlp._act(ws, len(self.mlp_dims) - 1, B)
        dcross, dmlp = self._head_bwd(ws, "mlp_linear.weight", xs[L], mlp_out, dlogits, B)
        dX = ws.mat("dX", B, D)
        self._mlp.bwd(ws, X, dmlp, B, train, dX, post_act_grad=True)
        dx0 = ws.mat("cn.dx0", B, D, zero=True)
        dx0.t[:B * D].zero_()
        dx = dcross
        for l in reversed(range(L)):
            xw = ws.mat(f"cn.xw{l}", B, 1)
            rt.ops.colsum(dx, B, D, rt.g(f"cn.b.{l}"))
            dxw = ws.mat(f"cn.dxw{l}", B, 1)
            rt.ops.cross_fuse_bwd(x0, xw, 1, dx, dx0, dxw, B, D)           # dx0 += dx*xw ; dxw = rowsum(dx*x0)
            tmp = ws.mat("cn.dxl", B, D)
            rt.ops.rowdot_bwd(xs[l], rt.w(f"cn.w.{l}.weight"), dxw, tmp, rt.g(f"cn.w.{l}.weight"), None, B, 1, D)
            rt.ops.add2d(tmp, dx, B, D, True)                              # dx <- dx + dxw (x) w_l
        rt.ops.add2d(dx, dX, B, D, True)
        rt.ops.add2d(dx0, dX, B, D, True)
        self._lin_bwd(ws, X, B, dX)
        return dX


# --------------------------------------------------------------------------------------------- DCNv2
class DCNv2(_CrossModel):
    def __init__(self, feature_dims, embed_dim, n_cross_layers, mlp_dims, dropout=0.2, model_structure="parallel",
                 use_low_rank_mixture=True, low_rank=32, num_experts=4,
                 l2_reg_embedding=1e-5, l2_reg_linear=1e-5, l2_reg_dnn=1e-5, l2_reg_cross=1e-5, config=None):
        super().__init__(feature_dims, embed_dim, l2_reg_embedding=l2_reg_embedding, l2_reg_linear=l2_reg_linear)
        self.model_name = 'dcnv2'
        self.model_structure = model_structure
        if model_structure not in ("stacked", "parallel"):
            # "crossnet_only" passes upstream's assert but crashes in its __init__ (self.dnn undefined)
            raise ValueError(f"model_structure={model_structure} not supported!")
        self.n_cross_layers, self.mlp_dims = n_cross_layers, tuple(mlp_dims)
        self.use_low_rank_mixture, self.low_rank, self.num_experts = bool(use_low_rank_mixture), low_rank, num_experts
        D = self.embed_output_dim
        if use_low_rank_mixture:
            self.crossnet = CrossNetMix(D, n_cross_layers, low_rank=low_rank, num_experts=num_experts)
        else:
            self.crossnet = CrossNetV2(D, n_cross_layers)
        self.dnn = MultiLayerPerceptron(D, mlp_dims, dropout, output_layer=False)
        final_dim = mlp_dims[-1] + (D if model_structure == "parallel" else 0)
        self.dnn_linear = nn.Linear(final_dim, 1, bias=False)
        self.output_layer = nn.Sigmoid()
        self.add_regularization_weight(self.reg_filter("dnn"), l2=l2_reg_dnn)
        self.add_regularization_weight(["dnn_linear.weight"], l2=l2_reg_linear)
        if use_low_rank_mixture:
            for lst in ("u_list", "v_list", "c_list"):
                self.add_regularization_weight([f"crossnet.{lst}.{l}" for l in range(n_cross_layers)], l2=l2_reg_cross)
        self._mlp_names, blk, bufs = mlp_group_names(["dnn"], self.dnn, "dnn")
        if use_low_rank_mixture:                                   # the shared gating vectors form one [n_exp, D] operand
            blk = blk + [("crossnet.G", [f"crossnet.gating.{e}.weight" for e in range(num_experts)])]
        self._finalize(blk, bufs, precision=precision_of(config), dropout=dropout)

    def _on_runtime_built(self):
        self._mlp = MlpGroup(self._rt, "dnn", 1, self.embed_output_dim, self.mlp_dims, self._mlp_names, bn=True, out_layer=False,
                             in_groups=None)

    # ---------------------------------------------------------------- CrossNetV2: x <- x0 * (x W^T) + b + x
    def _v2_tc(self) -> bool:
        """bf16 path: the D x D contractions of CrossNetV2 run on the tcgen05 GEMM (bf16 operands, fp32 accumulate and fp32 results);
        the Hadamard / bias / residual stages stay fp32.  TMA needs 16-byte aligned operand pitches."""
        return self._rt.bf16 and self.embed_output_dim % 8 == 0

    def _v2_operand(self, ws, l, cur: Mat, X: Mat | None, B) -> Mat:
        """bf16 GEMM operand of the layer-l input: the gathered bf16 embeddings themselves for l = 0, else a bf16 copy of x_l"""
        if l == 0 and X is not None and X.is_bf16:
            return X
        return self._rt.gemm_input(ws, f"cv2.xop{l}", cur, B, self.embed_output_dim)

    def _v2_fused(self, X: Mat | None) -> bool:
        """the layer's Hadamard / bias / residual stage runs inside the GEMM's epilogue (cdcmdr_gemm_bf16_t.cross_*): needs whole
        32-column chunks and the bf16 gathered embeddings as x0"""
        return self._v2_tc() and self.embed_output_dim % 32 == 0 and X is not None and X.is_bf16 and X.ld == self.embed_output_dim

    def _v2_fwd(self, ws, x0: Mat, B, X: Mat | None = None) -> Mat:
        rt, D = self._rt, self.embed_output_dim
        if self._v2_fused(X):
            # ONE launch per layer: the tcgen05 GEMM's epilogue applies x0 * acc + b + x (layer.py:339-343) on bf16 x0 / x boxes it
            # fetches by TMA and leaves the bf16 layer output (= the next layer's GEMM operand, reused by the weight-gradient GEMM of
            # the backward) and the raw fp32 product acc (the backward's dx0 += dx * acc) - no separate cast / Hadamard passes
            cur = X
            for l in range(self.n_cross_layers):
                wname = f"crossnet.w.{l}.weight"
                xw = ws.mat(f"cv2.xw{l}", B, D)
                nxt = ws.mat(f"cv2.xop{l + 1}", B, D, torch.bfloat16, zero=True)
                rt.ops.gemm_tc(A=cur.ptr, lda=cur.ld, a_rows=B, a_cols=D, a_mn=0, Bt=rt.Wb.data_ptr() + 2 * rt.o(wname), ldb=D,
                               b_rows=D, b_cols=D, b_mn=0, M=B, N=D, K=D, bias=rt.w(f"crossnet.b.{l}"), n_main=D,
                               out_main=nxt.ptr, ld_main=nxt.ld, out_aux=xw.ptr, ld_aux=xw.ld, cross_x0=X.ptr, cross_x=cur.ptr, ld_cross=D)
                cur = nxt
            return cur
        cur = x0
        for l in range(self.n_cross_layers):
            xw = ws.mat(f"cv2.xw{l}", B, D)
            wname = f"crossnet.w.{l}.weight"
            if self._v2_tc():
                cop = self._v2_operand(ws, l, cur, X, B)
                rt.ops.gemm_tc(A=cop.ptr, lda=cop.ld, a_rows=B, a_cols=D, a_mn=0, Bt=rt.Wb.data_ptr() + 2 * rt.o(wname), ldb=D,
                               b_rows=D, b_cols=D, b_mn=0, M=B, N=D, K=D, n_main=0, out_aux=xw.ptr, ld_aux=xw.ld)
            else:
                rt.ops.gemm_f32(A=cur.ptr, a_rs=cur.ld, a_cs=1, Bt=rt.w(wname), b_rs=D, b_cs=1, Cm=xw.ptr, c_rs=D,
                                M=B, N=D, K=D)
            nxt = ws.mat(f"cv2.x{l + 1}", B, D)
            rt.ops.cross_fuse_fwd(x0, cur, xw, D, rt.w(f"crossnet.b.{l}"), nxt, B, D)
            cur = nxt
        return cur

    def _v2_bwd(self, ws, x0: Mat, dout: Mat, B, X: Mat | None = None) -> Mat:
        """dout: fp32 [B, D] gradient of the cross output (consumed: updated in place) -> gradient w.r.t. x0."""
        rt, D, L = self._rt, self.embed_output_dim, self.n_cross_layers
        xs = [x0] + [ws.mat(f"cv2.x{l + 1}", B, D) for l in range(L)]
        dx0 = ws.mat("cv2.dx0", B, D)
        dx0.t[:B * D].zero_()
        dx = dout
        for l in reversed(range(L)):
            wname = f"crossnet.w.{l}.weight"
            W = rt.w(wname)
            xw = ws.mat(f"cv2.xw{l}", B, D)
            rt.ops.colsum(dx, B, D, rt.g(f"crossnet.b.{l}"))
            dxw = ws.mat("cv2.dxw", B, D)
            rt.ops.cross_fuse_bwd(x0, xw, D, dx, dx0, dxw, B, D)           # dx0 += dx*xw ; dxw = dx*x0
            if self._v2_tc():
                dop = rt.gemm_input(ws, "cv2.dxw_op", dxw, B, D)
                xop = self._v2_operand(ws, l, xs[l], X, B) if l == 0 else ws.mat(f"cv2.xop{l}", B, (D + 7) // 8 * 8, torch.bfloat16)
                # dW[n, k] = sum_b dxw[b, n] * x_l[b, k]: both operands are read as stored ([B, D], the batch is the reduction index)
                rt.ops.gemm_tc(A=dop.ptr, lda=dop.ld, a_rows=B, a_cols=D, a_mn=1, Bt=xop.ptr, ldb=xop.ld, b_rows=B, b_cols=D, b_mn=1,
                               M=D, N=D, K=B, n_main=0, out_aux=rt.g(wname), ld_aux=D, split_k="auto")
                # dx <- dx + dxw W   (W is [n, k]: the reduction index n is its row)
                rt.ops.gemm_tc(A=dop.ptr, lda=dop.ld, a_rows=B, a_cols=D, a_mn=0, Bt=rt.Wb.data_ptr() + 2 * rt.o(wname), ldb=D,
                               b_rows=D, b_cols=D, b_mn=1, M=B, N=D, K=D, n_main=0, out_aux=dx.ptr, ld_aux=dx.ld, accumulate=1)
                continue
            # dW[n, k] = sum_b dxw[b, n] * x_l[b, k]
            rt.ops.gemm_f32(A=dxw.ptr, a_rs=1, a_cs=D, Bt=xs[l].ptr, b_rs=1, b_cs=xs[l].ld, Cm=rt.g(wname), c_rs=D,
                            M=D, N=D, K=B, split_k=rt.ops.pick_split(D, D, 1, B))
            # dx <- dx + dxw W
            rt.ops.gemm_f32(A=dxw.ptr, a_rs=D, a_cs=1, Bt=W, b_rs=1, b_cs=D, Cm=dx.ptr, c_rs=dx.ld, M=B, N=D, K=D, accumulate=1)
        rt.ops.add2d(dx0, dx, B, D, True)
        return dx

    # ---------------------------------------------------------------- CrossNetMix
    def _mix_bufs(self, ws, l, B):
        ne, r, D = self.num_experts, self.low_rank, self.embed_output_dim
        return dict(g=ws.get(f"cmix.g{l}", (B, ne)), v1=ws.get(f"cmix.v1_{l}", (ne, B, r)), v2=ws.get(f"cmix.v2_{l}", (ne, B, r)),
                    u=ws.get(f"cmix.u{l}", (ne, B, D)))

    def _mix_tc(self) -> bool:
        """bf16 path: the contractions of CrossNetMix that involve the feature dimension D (x V_e, v U_e^T and their gradients) run on
        the tcgen05 GEMM with fp32 results; the r x r and gating products, tanh, softmax and the Hadamard stages stay fp32."""
        return self._rt.bf16 and self.embed_output_dim % 8 == 0 and self.low_rank % 8 == 0

    def _mix_fwd(self, ws, x0: Mat, B, X: Mat | None = None) -> Mat:
        rt, D, ne, r = self._rt, self.embed_output_dim, self.num_experts, self.low_rank
        ops = rt.ops
        tc = self._mix_tc()
        cur = x0
        for l in range(self.n_cross_layers):
            b = self._mix_bufs(ws, l, B)
            U, V, Cw = rt.w(f"crossnet.u_list.{l}"), rt.w(f"crossnet.v_list.{l}"), rt.w(f"crossnet.c_list.{l}")
            if tc:
                Ub = rt.Wb.data_ptr() + 2 * rt.o(f"crossnet.u_list.{l}")
                Vb = rt.Wb.data_ptr() + 2 * rt.o(f"crossnet.v_list.{l}")
            z = ws.get("cmix.z", (B, ne))
            ops.gemm_f32(A=cur.ptr, a_rs=cur.ld, a_cs=1, Bt=rt.w("crossnet.G"), b_rs=D, b_cs=1, Cm=z.data_ptr(), c_rs=ne, M=B, N=ne, K=D)
            ops.softmax_rows_fwd(z, ne, b["g"], ne, B, ne)
            # v1[e] = tanh(x V_e)            V_e stored [D, r]: Bt(n, k) = V_e[k, n]
            if tc:
                xop = self._v2_operand(ws, l, cur, X, B) if l == 0 else rt.gemm_input(ws, f"cmix.xop{l}", cur, B, D)
                for e in range(ne):
                    ops.gemm_tc(A=xop.ptr, lda=xop.ld, a_rows=B, a_cols=D, a_mn=0, Bt=Vb + 2 * e * D * r, ldb=r, b_rows=D, b_cols=r,
                                b_mn=1, M=B, N=r, K=D, n_main=0, out_aux=b["v1"].data_ptr() + 4 * e * B * r, ld_aux=r)
            else:
                ops.gemm_f32(A=cur.ptr, a_rs=cur.ld, a_cs=1, Bt=V, b_rs=1, b_cs=r, Cm=b["v1"].data_ptr(), c_rs=r, M=B, N=r, K=D,
                             G=ne, a_gs=0, b_gs=D * r, c_gs=B * r)
            ops.tanh_fwd(b["v1"], ne * B * r)
            # v2[e] = tanh(v1[e] C_e^T)      C_e stored [r, r]: Bt(n, k) = C_e[n, k]
            ops.gemm_f32(A=b["v1"].data_ptr(), a_rs=r, a_cs=1, Bt=Cw, b_rs=r, b_cs=1, Cm=b["v2"].data_ptr(), c_rs=r, M=B, N=r, K=r,
                         G=ne, a_gs=B * r, b_gs=r * r, c_gs=B * r)
            ops.tanh_fwd(b["v2"], ne * B * r)
            # u[e] = v2[e] U_e^T             U_e stored [D, r]: Bt(n, k) = U_e[n, k]
            if tc:
                v2op = rt.gemm_input(ws, f"cmix.v2op{l}", Mat(b["v2"], 0, r), ne * B, r)
                for e in range(ne):
                    ops.gemm_tc(A=v2op.ptr + 2 * e * B * v2op.ld, lda=v2op.ld, a_rows=B, a_cols=r, a_mn=0, Bt=Ub + 2 * e * D * r, ldb=r,
                                b_rows=D, b_cols=r, b_mn=0, M=B, N=D, K=r, n_main=0, out_aux=b["u"].data_ptr() + 4 * e * B * D, ld_aux=D)
            else:
                ops.gemm_f32(A=b["v2"].data_ptr(), a_rs=r, a_cs=1, Bt=U, b_rs=r, b_cs=1, Cm=b["u"].data_ptr(), c_rs=D, M=B, N=D, K=r,
                             G=ne, a_gs=B * r, b_gs=D * r, c_gs=B * D)
            nxt = ws.mat(f"cmix.x{l + 1}", B, D)
            ops.crossmix_combine_fwd(x0, cur, b["u"], b["g"], rt.w(f"crossnet.bias.{l}"), nxt, B, D, ne)
            cur = nxt
        return cur

    def _mix_bwd(self, ws, x0: Mat, dout: Mat, B, X: Mat | None = None) -> Mat:
        rt, D, ne, r, L = self._rt, self.embed_output_dim, self.num_experts, self.low_rank, self.n_cross_layers
        ops = rt.ops
        tc = self._mix_tc()
        ldp = (D + 7) // 8 * 8
        xs = [x0] + [ws.mat(f"cmix.x{l + 1}", B, D) for l in range(L)]
        dx0 = ws.mat("cmix.dx0", B, D)
        dx0.t[:B * D].zero_()
        du = ws.get("cmix.du", (ne, B, D))
        dv2 = ws.get("cmix.dv2", (ne, B, r))
        dv1 = ws.get("cmix.dv1", (ne, B, r))
        dgate = ws.get("cmix.dgate", (B, ne))
        dz = ws.get("cmix.dz", (B, ne))
        dx = dout
        first = True
        for l in reversed(range(L)):
            b = self._mix_bufs(ws, l, B)
            x = xs[l]
            U, V, Cw = rt.w(f"crossnet.u_list.{l}"), rt.w(f"crossnet.v_list.{l}"), rt.w(f"crossnet.c_list.{l}")
            gU, gV, gC = rt.g(f"crossnet.u_list.{l}"), rt.g(f"crossnet.v_list.{l}"), rt.g(f"crossnet.c_list.{l}")
            ops.crossmix_combine_bwd(x0, b["u"], b["g"], rt.w(f"crossnet.bias.{l}"), dx, du, dgate, dx0, B, D, ne)
            ops.colsum(Mat(du, 0, D), ne * B, D, rt.g(f"crossnet.bias.{l}"))
            sp = ops.pick_split(D, r, ne, B)
            if tc:
                Ub = rt.Wb.data_ptr() + 2 * rt.o(f"crossnet.u_list.{l}")
                Vb = rt.Wb.data_ptr() + 2 * rt.o(f"crossnet.v_list.{l}")
                duop = rt.gemm_input(ws, "cmix.duop", Mat(du, 0, D), ne * B, D)
                v2op = ws.mat(f"cmix.v2op{l}", ne * B, r, torch.bfloat16)          # made by the forward
                for e in range(ne):
                    du_e = duop.ptr + 2 * e * B * duop.ld
                    # dU_e[d, j] = sum_b du[e][b, d] * v2[e][b, j]     (both read as stored: the batch is the reduction index)
                    ops.gemm_tc(A=du_e, lda=duop.ld, a_rows=B, a_cols=D, a_mn=1, Bt=v2op.ptr + 2 * e * B * r, ldb=r, b_rows=B, b_cols=r,
                                b_mn=1, M=D, N=r, K=B, n_main=0, out_aux=gU + 4 * e * D * r, ld_aux=r, split_k="auto")
                    # dv2[e] = du[e] U_e                               Bt(n=j, k=d) = U_e[d, j]
                    ops.gemm_tc(A=du_e, lda=duop.ld, a_rows=B, a_cols=D, a_mn=0, Bt=Ub + 2 * e * D * r, ldb=r, b_rows=D, b_cols=r,
                                b_mn=1, M=B, N=r, K=D, n_main=0, out_aux=dv2.data_ptr() + 4 * e * B * r, ld_aux=r)
            else:
                # dU_e[d, j] = sum_b du[e][b, d] * v2[e][b, j]
                ops.gemm_f32(A=du.data_ptr(), a_rs=1, a_cs=D, Bt=b["v2"].data_ptr(), b_rs=1, b_cs=r, Cm=gU, c_rs=r, M=D, N=r, K=B,
                             G=ne, a_gs=B * D, b_gs=B * r, c_gs=D * r, split_k=sp)
                # dv2[e] = (du[e] U_e) * (1 - v2^2)      Bt(n=j, k=d) = U_e[d, j]
                ops.gemm_f32(A=du.data_ptr(), a_rs=D, a_cs=1, Bt=U, b_rs=1, b_cs=r, Cm=dv2.data_ptr(), c_rs=r, M=B, N=r, K=D,
                             G=ne, a_gs=B * D, b_gs=D * r, c_gs=B * r)
            ops.tanh_bwd(b["v2"], dv2, ne * B * r)
            # dC_e[i, j] = sum_b dv2[e][b, i] * v1[e][b, j]
            ops.gemm_f32(A=dv2.data_ptr(), a_rs=1, a_cs=r, Bt=b["v1"].data_ptr(), b_rs=1, b_cs=r, Cm=gC, c_rs=r, M=r, N=r, K=B,
                         G=ne, a_gs=B * r, b_gs=B * r, c_gs=r * r, split_k=ops.pick_split(r, r, ne, B))
            # dv1[e] = (dv2[e] C_e) * (1 - v1^2)     Bt(n=j, k=i) = C_e[i, j]
            ops.gemm_f32(A=dv2.data_ptr(), a_rs=r, a_cs=1, Bt=Cw, b_rs=1, b_cs=r, Cm=dv1.data_ptr(), c_rs=r, M=B, N=r, K=r,
                         G=ne, a_gs=B * r, b_gs=r * r, c_gs=B * r)
            ops.tanh_bwd(b["v1"], dv1, ne * B * r)
            # dV_e[d, j] = sum_b x[b, d] * dv1[e][b, j]
            if tc:
                xop = self._v2_operand(ws, l, x, X, B) if l == 0 else ws.mat(f"cmix.xop{l}", B, ldp, torch.bfloat16)
                dv1op = rt.gemm_input(ws, "cmix.dv1op", Mat(dv1, 0, r), ne * B, r)
                for e in range(ne):
                    ops.gemm_tc(A=xop.ptr, lda=xop.ld, a_rows=B, a_cols=D, a_mn=1, Bt=dv1op.ptr + 2 * e * B * r, ldb=r, b_rows=B,
                                b_cols=r, b_mn=1, M=D, N=r, K=B, n_main=0, out_aux=gV + 4 * e * D * r, ld_aux=r, split_k="auto")
            else:
                ops.gemm_f32(A=x.ptr, a_rs=1, a_cs=x.ld, Bt=dv1.data_ptr(), b_rs=1, b_cs=r, Cm=gV, c_rs=r, M=D, N=r, K=B,
                             G=ne, a_gs=0, b_gs=B * r, c_gs=D * r, split_k=sp)
            ops.softmax_rows_bwd(b["g"], ne, dgate, ne, dz, ne, B, ne)
            # dG[e, d] (+)= sum_b dz[b, e] * x[b, d]   (the gating vectors are shared by every layer)
            ops.gemm_f32(A=dz.data_ptr(), a_rs=1, a_cs=ne, Bt=x.ptr, b_rs=1, b_cs=x.ld, Cm=rt.g("crossnet.G"), c_rs=D, M=ne, N=D, K=B,
                         accumulate=0 if first else 1, split_k=ops.pick_split(ne, D, 1, B))     # 4 x D outputs: the batch must be split
            first = False
            # dx <- dx + sum_e dv1[e] V_e^T + dz G
            for e in range(ne):
                if tc:
                    ops.gemm_tc(A=dv1op.ptr + 2 * e * B * r, lda=r, a_rows=B, a_cols=r, a_mn=0, Bt=Vb + 2 * e * D * r, ldb=r, b_rows=D,
                                b_cols=r, b_mn=0, M=B, N=D, K=r, n_main=0, out_aux=dx.ptr, ld_aux=dx.ld, accumulate=1)
                    continue
                ops.gemm_f32(A=dv1.data_ptr() + 4 * e * B * r, a_rs=r, a_cs=1, Bt=V + 4 * e * D * r, b_rs=r, b_cs=1, Cm=dx.ptr, c_rs=dx.ld,
                             M=B, N=D, K=r, accumulate=1)
            ops.gemm_f32(A=dz.data_ptr(), a_rs=ne, a_cs=1, Bt=rt.w("crossnet.G"), b_rs=1, b_cs=D, Cm=dx.ptr, c_rs=dx.ld, M=B, N=D, K=ne,
                         accumulate=1)
        ops.add2d(dx0, dx, B, D, True)
        return dx

    # ---------------------------------------------------------------- model program
    def _program_fwd(self, ws, X: Mat, B, train):
        rt, D = self._rt, self.embed_output_dim
        x0 = self._x32(ws, X, B)
        cross = self._mix_fwd(ws, x0, B, X) if self.use_low_rank_mixture else self._v2_fwd(ws, x0, B, X)
        if self.model_structure == "parallel":
            mlp_out = self._mlp.fwd(ws, X, B, train)
            logit = self._head_fwd(ws, "dnn_linear.weight", cross, mlp_out, B)
        else:
            mlp_in = rt.gemm_input(ws, "dnn.in_op", cross, B, D)
            mlp_out = self._mlp.fwd(ws, mlp_in, B, train)
            logit = self._head_fwd(ws, "dnn_linear.weight", None, mlp_out, B)
        return logit, self._lin_fwd(ws, X, B)

    def _cross_out(self, ws, B, X: Mat | None = None) -> Mat:
        tag = "cmix" if self.use_low_rank_mixture else "cv2"
        if not self.use_low_rank_mixture and self._v2_fused(X):     # the fused layers leave bf16 outputs (_v2_fwd)
            return ws.mat(f"cv2.xop{self.n_cross_layers}", B, self.embed_output_dim, torch.bfloat16)
        return ws.mat(f"{tag}.x{self.n_cross_layers}", B, self.embed_output_dim)

    def _program_bwd(self, ws, X: Mat, B, train, dlogits: Mat):
        rt, D = self._rt, self.embed_output_dim
        x0 = ws.mat("X32", B, D) if X.is_bf16 else X
        cross = self._cross_out(ws, B, X)
        mlp_out = self._mlp._act(ws, len(self.mlp_dims) - 1, B)
        cross_bwd = (lambda w, a, d, n: (self._mix_bwd if self.use_low_rank_mixture else self._v2_bwd)(w, a, d, n, X))
        dX = ws.mat("dX", B, D)
        if self.model_structure == "parallel":
            dcross, dmlp = self._head_bwd(ws, "dnn_linear.weight", cross, mlp_out, dlogits, B)
            self._mlp.bwd(ws, X, dmlp, B, train, dX, post_act_grad=True)
            dx = cross_bwd(ws, x0, dcross, B)
            rt.ops.add2d(dx, dX, B, D, True)
        else:
            _, dmlp = self._head_bwd(ws, "dnn_linear.weight", None, mlp_out, dlogits, B)
            mlp_in = rt.gemm_input(ws, "dnn.in_op", cross, B, D) if rt.bf16 else cross
            dcross = ws.mat("dnn.dX", B, D)
            self._mlp.bwd(ws, mlp_in, dmlp, B, train, dcross, post_act_grad=True)
            dx = cross_bwd(ws, x0, dcross, B)
            rt.ops.add2d(dx, dX, B, D, False)
        self._lin_bwd(ws, X, B, dX)
        return dX
