"""Per-model runtime: the flat parameter arena (weights / gradients / Adam moments / per-element L2 coefficients),
the BatchNorm buffer arena, workspaces, and the building blocks every model program is made of
(grouped Linear forward / weight-gradient / input-gradient and the MLP groups of the reference,
model/layer.py:178-206).  All arithmetic happens in libcdcmdr.so; this file only sequences launches.
"""
from __future__ import annotations

import torch

from .core import Mat, Ops, Workspace


class Arena:
    """Flat fp32 storage for all dense parameters of one model.  Parameters that one GEMM consumes together are
    placed contiguously (the model passes the order), so packing weights for a concatenated-N or grouped GEMM is
    a pointer offset, and the weight-gradient GEMM writes straight into the gradient arena."""

    def __init__(self):
        self.off = {}
        self.shape = {}
        self.n = 0

    def add(self, name, shape):
        if name in self.off:
            raise KeyError(f"duplicate arena entry {name}")
        n = 1
        for s in shape:
            n *= int(s)
        self.off[name] = self.n
        self.shape[name] = tuple(int(s) for s in shape)
        self.n += n
        self.n = (self.n + 3) & ~3          # keep every entry 16-byte aligned for vector loads

    def numel(self, name):
        n = 1
        for s in self.shape[name]:
            n *= s
        return n


class Runtime:
    def __init__(self, device, precision="fp32"):
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self.precision = precision
        self.device = torch.device(device)
        self.ops = Ops(self.device)
        self.params = Arena()          # trainable dense parameters
        self.buffers = Arena()         # BatchNorm running statistics
        self.W = self.G = self.M = self.V = self.L2 = self.present = self.Bf = None
        self._ws = {}
        self.step_state = None
        self.dropout = 0.0
        self.base_seed = 2000
        self._salt = 0

    # ---------------------------------------------------------------- storage
    def allocate(self):
        dev = self.device
        self.W = torch.zeros(max(self.params.n, 4), dtype=torch.float32, device=dev)
        self.G = torch.zeros_like(self.W)
        self.L2 = torch.zeros_like(self.W)
        self.present = torch.ones(self.W.numel(), dtype=torch.uint8, device=dev)
        self.Bf = torch.zeros(max(self.buffers.n, 4), dtype=torch.float32, device=dev)
        self.M = self.V = None
        self._ws = {}
        self.ops = Ops(dev)
        self.step_state = None

    def ensure_opt_state(self):
        if self.M is None:
            self.M = torch.zeros_like(self.W)
            self.V = torch.zeros_like(self.W)
        if self.step_state is None:
            self.step_state = self.ops.step_state_new()

    def view(self, name) -> torch.Tensor:
        o = self.params.off[name]
        return self.W[o:o + self.params.numel(name)].view(self.params.shape[name])

    def grad_view(self, name) -> torch.Tensor:
        o = self.params.off[name]
        return self.G[o:o + self.params.numel(name)].view(self.params.shape[name])

    def buf_view(self, name) -> torch.Tensor:
        o = self.buffers.off[name]
        return self.Bf[o:o + self.buffers.numel(name)].view(self.buffers.shape[name])

    def w(self, name, extra=0) -> int:
        return self.W.data_ptr() + 4 * (self.params.off[name] + extra)

    def g(self, name, extra=0) -> int:
        return self.G.data_ptr() + 4 * (self.params.off[name] + extra)

    def b(self, name, extra=0) -> int:
        return self.Bf.data_ptr() + 4 * (self.buffers.off[name] + extra)

    def ws(self, B) -> Workspace:
        w = self._ws.get(B)
        if w is None:
            w = Workspace(self.device)
            self._ws[B] = w
        return w

    def next_salt(self) -> int:
        self._salt += 1
        return self._salt

    @property
    def seed_ptr(self):
        if self.step_state is None:
            self.step_state = self.ops.step_state_new()
        return Ops.seed_ptr(self.step_state)

    # ---------------------------------------------------------------- grouped Linear building blocks (fp32 path)
    def lin_fwd(self, X: Mat, K, w_addr, N, b_addr, Y: Mat, M, *, G=1, x_gs=0, w_gs=None, b_gs=None, y_gs=None,
                relu=False, drop=0.0, salt=0):
        """Y[g] = act(X[g] @ W[g]^T + b[g]);  W[g] is [N, K] row-major at w_addr + g*w_gs floats."""
        self.ops.gemm_f32(A=X.ptr, a_rs=X.ld, a_cs=1, Bt=w_addr, b_rs=K, b_cs=1, Cm=Y.ptr, c_rs=Y.ld, M=M, N=N, K=K, G=G,
                          a_gs=x_gs, b_gs=N * K if w_gs is None else w_gs, c_gs=N if y_gs is None else y_gs,
                          bias=b_addr, bias_gs=N if b_gs is None else b_gs, act=1 if relu else 0,
                          drop_p=drop, seed_ptr=self.seed_ptr if drop > 0 else None, salt=salt)

    def lin_bwd_w(self, dY: Mat, X: Mat, K, gw_addr, N, M, *, G=1, dy_gs=None, x_gs=0, w_gs=None):
        """dW[g][n, k] = sum_m dY[g][m, n] * X[g][m, k]  -> gradient arena."""
        split = Ops.pick_split(N, K, G, M)
        self.ops.gemm_f32(A=dY.ptr, a_rs=1, a_cs=dY.ld, Bt=X.ptr, b_rs=1, b_cs=X.ld, Cm=gw_addr, c_rs=K, M=N, N=K, K=M, G=G,
                          a_gs=N if dy_gs is None else dy_gs, b_gs=x_gs, c_gs=N * K if w_gs is None else w_gs, split_k=split)

    def lin_bwd_x(self, dY: Mat, K, w_addr, N, dX: Mat, M, *, G=1, dy_gs=None, w_gs=None, dx_gs=0, mask: Mat | None = None,
                  mask_gs=0, mask_scale=1.0, accumulate=False):
        """dX[g][m, k] (+)= sum_n dY[g][m, n] * W[g][n, k]; optional ReLU/dropout mask from the forward activation."""
        self.ops.gemm_f32(A=dY.ptr, a_rs=dY.ld, a_cs=1, Bt=w_addr, b_rs=1, b_cs=K, Cm=dX.ptr, c_rs=dX.ld, M=M, N=K, K=N, G=G,
                          a_gs=N if dy_gs is None else dy_gs, b_gs=N * K if w_gs is None else w_gs, c_gs=dx_gs,
                          mask=mask.ptr if mask is not None else None, mask_rs=mask.ld if mask is not None else 0,
                          mask_gs=mask_gs, mask_scale=mask_scale, accumulate=1 if accumulate else 0)


class MlpGroup:
    """G parallel MLPs with identical layer sizes (reference MultiLayerPerceptron, model/layer.py:178-206):
    per layer Linear -> [BatchNorm1d] -> ReLU -> Dropout, optional final Linear(., 1).

    names: dict with arena entry names:  W[j], b[j] ([G, d_j, d_{j-1}] / [G, d_j]); with bn: gamma[j], beta[j] and buffer
    names rmean[j], rvar[j]; with out_layer: Wout ([G, d_last]), bout ([G]).
    in_groups: layer-0 input mapping: list of (input_block, e_lo, e_hi): MLPs e_lo..e_hi-1 read columns
    [input_block*in_dim, (input_block+1)*in_dim) of X (one concatenated-N GEMM per entry); None: MLP g reads block g.
    """

    def __init__(self, rt: Runtime, tag, G, in_dim, dims, names, *, bn, out_layer, in_groups):
        self.rt, self.tag, self.G, self.in_dim, self.dims = rt, tag, G, in_dim, tuple(dims)
        self.names, self.bn, self.out_layer, self.in_groups = names, bn, out_layer, in_groups
        self.salts = [rt.next_salt() for _ in dims]

    # activation dtype of the fp32 path
    def _act(self, ws: Workspace, j, B) -> Mat:
        return ws.mat(f"{self.tag}.A{j}", B, self.G * self.dims[j])

    def fwd(self, ws: Workspace, X: Mat, B, train) -> Mat:
        rt, G = self.rt, self.G
        use_bn = self.bn and B != 1                              # layer.py:202-204
        drop = rt.dropout if train else 0.0
        prev, prev_d = X, self.in_dim
        for j, d in enumerate(self.dims):
            fused_act = not use_bn
            Y = self._act(ws, j, B) if fused_act else ws.mat(f"{self.tag}.Z{j}", B, G * d)
            if j == 0 and self.in_groups is None:               # MLP g reads input block g: one grouped launch
                rt.lin_fwd(prev, prev_d, rt.w(self.names["W"][0]), d, rt.w(self.names["b"][0]), Y, B, G=G, x_gs=prev_d,
                           relu=fused_act, drop=drop if fused_act else 0.0, salt=self.salts[j])
            elif j == 0:
                for (blk, e0, e1) in self.in_groups:
                    rt.lin_fwd(prev.cols(blk * prev_d), prev_d, rt.w(self.names["W"][0], e0 * d * prev_d), (e1 - e0) * d,
                               rt.w(self.names["b"][0], e0 * d), Y.cols(e0 * d), B, relu=fused_act,
                               drop=drop if fused_act else 0.0, salt=self.salts[j] + 7919 * e0)
            else:
                rt.lin_fwd(prev, prev_d, rt.w(self.names["W"][j]), d, rt.w(self.names["b"][j]), Y, B, G=G, x_gs=prev_d,
                           relu=fused_act, drop=drop if fused_act else 0.0, salt=self.salts[j])
            if use_bn:
                A = self._act(ws, j, B)
                sm = ws.get(f"{self.tag}.bnsave{j}", (2, G * d))
                desc = rt.ops.bn_desc(rt.w(self.names["gamma"][j]), rt.w(self.names["beta"][j]),
                                      rt.b(self.names["rmean"][j]), rt.b(self.names["rvar"][j]),
                                      sm.data_ptr(), sm.data_ptr() + 4 * G * d, train, True,
                                      drop_p=drop, seed_ptr=rt.seed_ptr if drop > 0 else None, salt=self.salts[j])
                rt.ops.bn_fwd(desc, Y, A, B, G * d)
                Y = A
            prev, prev_d = Y, d
        if self.out_layer:
            L = ws.mat(f"{self.tag}.logit", B, G)
            # logit[b, g] = A_last[b, g*d:(g+1)*d] . Wout[g] + bout[g]
            rt.ops.gemm_f32(A=prev.ptr, a_rs=prev.ld, a_cs=1, Bt=rt.w(self.names["Wout"]), b_rs=prev_d, b_cs=1, Cm=L.ptr,
                            c_rs=G, M=B, N=1, K=prev_d, G=G, a_gs=prev_d, b_gs=prev_d, c_gs=1,
                            bias=rt.w(self.names["bout"]), bias_gs=1)
            return L
        return prev

    def fwd_layer0_only(self, ws: Workspace, X: Mat, B):
        """The first (concatenated-N) layer alone, with the training epilogue: used by bench.py to time the dominant GEMM."""
        rt, d = self.rt, self.dims[0]
        fused_act = not self.bn
        Y = self._act(ws, 0, B) if fused_act else ws.mat(f"{self.tag}.Z0", B, self.G * d)
        for (blk, e0, e1) in (self.in_groups or [(0, 0, self.G)]):
            rt.lin_fwd(X.cols(blk * self.in_dim), self.in_dim, rt.w(self.names["W"][0], e0 * d * self.in_dim), (e1 - e0) * d,
                       rt.w(self.names["b"][0], e0 * d), Y.cols(e0 * d), B, relu=fused_act,
                       drop=rt.dropout if fused_act else 0.0, salt=self.salts[0] + 7919 * e0)

    def bwd(self, ws: Workspace, X: Mat, dOut: Mat, B, train, dX: Mat | None, accumulate=False):
        """dOut: gradient w.r.t. the group's output: dlogits [B, G] (out_layer), else the gradient of the last
        post-activation [B, G*d_last] (bn) or of the last PRE-activation (no bn: the caller applied the ReLU mask)."""
        rt, G, nl = self.rt, self.G, len(self.dims)
        use_bn = self.bn and B != 1
        drop = rt.dropout if train else 0.0
        keep = 1.0 / (1.0 - drop) if drop > 0 else 1.0
        d_last = self.dims[-1]
        cur = dOut
        if self.out_layer:
            A_last = self._act(ws, nl - 1, B)
            # dWout[g, k] = sum_b dlogit[b, g] * A_last[b, g*d + k]
            split = rt.ops.pick_split(1, d_last, G, B)
            rt.ops.gemm_f32(A=dOut.ptr, a_rs=0, a_cs=dOut.ld, Bt=A_last.ptr, b_rs=1, b_cs=A_last.ld, Cm=rt.g(self.names["Wout"]),
                            c_rs=d_last, M=1, N=d_last, K=B, G=G, a_gs=1, b_gs=d_last, c_gs=d_last, split_k=split)
            rt.ops.colsum(dOut, B, G, rt.g(self.names["bout"]))
            dA = ws.mat(f"{self.tag}.dA{nl - 1}", B, G * d_last)
            # dA_last[b, g*d + k] = dlogit[b, g] * Wout[g, k]    (ReLU mask applied below / by bn_bwd)
            mask = None if use_bn else A_last
            rt.ops.gemm_f32(A=dOut.ptr, a_rs=dOut.ld, a_cs=1, Bt=rt.w(self.names["Wout"]), b_rs=1, b_cs=1, Cm=dA.ptr,
                            c_rs=dA.ld, M=B, N=d_last, K=1, G=G, a_gs=1, b_gs=d_last, c_gs=d_last,
                            mask=mask.ptr if mask is not None else None, mask_rs=mask.ld if mask is not None else 0,
                            mask_gs=d_last, mask_scale=keep)
            cur = dA
        for j in reversed(range(nl)):
            d = self.dims[j]
            prev_d = self.in_dim if j == 0 else self.dims[j - 1]
            if use_bn:
                Z = ws.mat(f"{self.tag}.Z{j}", B, G * d)
                A = self._act(ws, j, B)
                sm = ws.get(f"{self.tag}.bnsave{j}", (2, G * d))
                dZ = ws.mat(f"{self.tag}.dZ{j}", B, G * d)
                desc = rt.ops.bn_desc(rt.w(self.names["gamma"][j]), rt.w(self.names["beta"][j]), None, None,
                                      sm.data_ptr(), sm.data_ptr() + 4 * G * d, train, True, drop_p=drop,
                                      seed_ptr=rt.seed_ptr if drop > 0 else None)
                rt.ops.bn_bwd(desc, Z, A, cur, dZ, rt.g(self.names["gamma"][j]), rt.g(self.names["beta"][j]), False, B, G * d)
                cur = dZ
            # cur is now dZ_j  [B, G*d]
            rt.ops.colsum(cur, B, G * d, rt.g(self.names["b"][j]))
            if j == 0 and self.in_groups is None:
                rt.lin_bwd_w(cur, X, prev_d, rt.g(self.names["W"][0]), d, B, G=G, x_gs=prev_d)
                if dX is not None:
                    rt.lin_bwd_x(cur, prev_d, rt.w(self.names["W"][0]), d, dX, B, G=G, dx_gs=prev_d, accumulate=accumulate)
            elif j == 0:
                for (blk, e0, e1) in self.in_groups:
                    rt.lin_bwd_w(cur.cols(e0 * d), X.cols(blk * prev_d), prev_d, rt.g(self.names["W"][0], e0 * d * prev_d),
                                 (e1 - e0) * d, B)
                if dX is not None:
                    seen = set()
                    for (blk, e0, e1) in self.in_groups:
                        acc = accumulate or (blk in seen)
                        seen.add(blk)
                        rt.lin_bwd_x(cur.cols(e0 * d), prev_d, rt.w(self.names["W"][0], e0 * d * prev_d), (e1 - e0) * d,
                                     dX.cols(blk * prev_d), B, accumulate=acc)
            else:
                A_prev = self._act(ws, j - 1, B)
                rt.lin_bwd_w(cur, A_prev, prev_d, rt.g(self.names["W"][j]), d, B, G=G, x_gs=prev_d)
                dA = ws.mat(f"{self.tag}.dA{j - 1}", B, G * prev_d)
                mask = None if use_bn else A_prev
                rt.lin_bwd_x(cur, prev_d, rt.w(self.names["W"][j]), d, dA, B, G=G, dx_gs=prev_d, mask=mask, mask_gs=prev_d,
                             mask_scale=keep)
                cur = dA
