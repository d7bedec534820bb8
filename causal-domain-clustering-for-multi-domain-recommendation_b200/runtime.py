"""Per-model runtime: the flat parameter arena (weights / gradients / Adam moments / per-element L2 coefficients, plus
the bf16 operand copy of the weights on the tensor-core path), the BatchNorm buffer arena, workspaces, and the building
blocks every model program is made of (grouped Linear forward / weight-gradient / input-gradient and the MLP groups
of the reference, model/layer.py:178-206).  All arithmetic happens in libcdcmdr.so; this file only sequences launches.

Two precisions share the same programs:
  fp32  every operand fp32, GEMMs on CUDA cores (cdcmdr_gemm_f32): the exact-parity path (<= 1e-4 vs the reference);
  bf16  activations / activation gradients / weight operands bf16, GEMMs on tcgen05 tensor cores with fp32 TMEM
        accumulation (cdcmdr_gemm_bf16_tc); master weights, gradients, Adam, BatchNorm statistics, gate logits, softmax,
        loss stay fp32 (<= 2e-2 vs the reference).
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
        self.n = (self.n + 7) & ~7          # every block 32-byte aligned in fp32 / 16-byte aligned in the bf16 copy (TMA)

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
        self.W = self.G = self.M = self.V = self.L2 = self.present = self.Bf = self.Wb = None
        self._ws = {}
        self.step_state = None
        self.dropout = 0.0
        self.base_seed = 2000
        self._salt = 0
        self.dp = None                 # parallel.DataParallel when the model is one replica of a data-parallel group

    @property
    def bf16(self) -> bool:
        return self.precision == "bf16"

    @property
    def act_dtype(self):
        return torch.bfloat16 if self.bf16 else torch.float32

    # ---------------------------------------------------------------- storage
    def allocate(self):
        dev = self.device
        self.W = torch.zeros(max(self.params.n, 8), dtype=torch.float32, device=dev)
        self.G = torch.zeros_like(self.W)
        self.L2 = torch.zeros_like(self.W)
        self.present = torch.ones(self.W.numel(), dtype=torch.uint8, device=dev)
        self.Bf = torch.zeros(max(self.buffers.n, 8), dtype=torch.float32, device=dev)
        self.Wb = torch.zeros(self.W.numel(), dtype=torch.bfloat16, device=dev) if self.bf16 else None
        self.M = self.V = None
        self._ws = {}
        self.ops = Ops(dev)
        self.step_state = None

    def refresh_operands(self):
        """bf16 path: re-derive the bf16 GEMM operand copy of the (fp32 master) weights; one cast over the arena."""
        if self.bf16:
            n = self.W.numel()
            self.ops.cast_f32_bf16(Mat(self.W, 0, n), Mat(self.Wb, 0, n), 1, n)

    def ensure_opt_state(self):
        if self.M is None:
            self.M = torch.zeros_like(self.W)
            self.V = torch.zeros_like(self.W)
        if self.step_state is None:
            self.step_state = self.ops.step_state_new()

    def view(self, name) -> torch.Tensor:
        o = self.params.off[name]
        return self.W[o:o + self.params.numel(name)].view(self.params.shape[name])

    def buf_view(self, name) -> torch.Tensor:
        o = self.buffers.off[name]
        return self.Bf[o:o + self.buffers.numel(name)].view(self.buffers.shape[name])

    def o(self, name, extra=0) -> int:
        """element offset of a parameter (block) in the arena"""
        return self.params.off[name] + extra

    def w(self, name, extra=0) -> int:
        return self.W.data_ptr() + 4 * (self.params.off[name] + extra)

    def g(self, name, extra=0) -> int:
        return self.G.data_ptr() + 4 * (self.params.off[name] + extra)

    def b(self, name, extra=0) -> int:
        return self.Bf.data_ptr() + 4 * (self.buffers.off[name] + extra)

    def side_stream(self):
        """Second CUDA stream for the step's independent branches (None on a non-CUDA device, i.e. under the host emulator)."""
        if self.device.type != "cuda":
            return None
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(self.device)
        return self._side

    def mark_input_grad(self):
        """Called by a model program when a gradient w.r.t. the gathered embeddings has just been enqueued.  While a fused step
        is armed (`arm_input_grad`) this records an event: the program promises that the gradient is FINAL at this point, so the
        embedding backward may start on the side stream at that event, next to the rest of the dense backward."""
        if getattr(self, "_dx_armed", False) and self.device.type == "cuda":
            self._dx_event = torch.cuda.Event()
            self._dx_event.record(torch.cuda.current_stream(self.device))

    def arm_input_grad(self, on: bool):
        self._dx_armed, self._dx_event = bool(on), None

    # Workspaces are keyed by the exact batch size (a captured CUDA graph holds their addresses), but only the most recently used
    # ones are kept: the CDC probing loop (run.py:528-594) concatenates 1..7 per-domain batches whose row total changes whenever a
    # loader yields its short last batch, and every distinct total would otherwise pin a full activation workspace for good.
    WS_MAX_ENTRIES = 12
    WS_MAX_BYTES = 32 << 30

    def ws(self, B) -> Workspace:
        w = self._ws.pop(B, None)
        if w is None:
            w = Workspace(self.device)
        self._ws[B] = w                                  # most recently used last
        self._evict_workspaces(keep=B)
        return w

    def pin_ws(self, B):
        """A CUDA graph was captured over the workspace of batch size B: it must never be evicted."""
        self._ws_pinned = getattr(self, "_ws_pinned", set()) | {B}
        self.ops.graph_pinned = True                     # scratch buffers that grow later keep their old storage alive

    def _evict_workspaces(self, keep):
        pinned = getattr(self, "_ws_pinned", ())
        def total():
            return sum(w.nbytes() for w in self._ws.values())
        for key in list(self._ws):
            if len(self._ws) <= self.WS_MAX_ENTRIES and total() <= self.WS_MAX_BYTES:
                break
            if key == keep or key in pinned:
                continue
            del self._ws[key]                            # back to torch's caching allocator (stream-ordered reuse)

    def ws_bytes(self) -> int:
        return sum(w.nbytes() for w in self._ws.values())

    def next_salt(self) -> int:
        self._salt += 1
        return self._salt

    @property
    def seed_ptr(self):
        if self.step_state is None:
            self.step_state = self.ops.step_state_new()
        return Ops.seed_ptr(self.step_state)

    def gemm_input(self, ws: Workspace, name, m: Mat, rows, cols) -> Mat:
        """A GEMM operand view of an fp32 matrix: itself on the fp32 path, a bf16 copy (ld padded to 8) on the bf16 path."""
        if not self.bf16 or m.is_bf16:
            return m
        ld = (cols + 7) // 8 * 8
        out = ws.mat(name, rows, ld, torch.bfloat16, zero=True)
        self.ops.cast_f32_bf16(m, out, rows, cols)
        return out

    def batch_rows(self, B) -> int:
        """rows of the single-device batch behind this rank's B rows: BatchNorm is skipped on a ONE-row batch (layer.py:202-204) -
        under data-parallel replicas that is a statement about the global batch"""
        return self.dp.global_rows(B) if self.dp is not None else B

    # ---------------------------------------------------------------- BatchNorm (cross-replica when data-parallel)
    def bn_fwd(self, desc, Z: Mat, A: Mat, B, Cn):
        if self.dp is not None and desc.train:
            self.ops.bn_fwd_sync(desc, Z, A, B, Cn, self.dp.global_rows(B), self.dp.all_reduce_sum)
        else:
            self.ops.bn_fwd(desc, Z, A, B, Cn)

    def bn_bwd(self, desc, Z: Mat, A, dA: Mat, dZ: Mat, dgamma, dbeta, accumulate, B, Cn):
        if self.dp is not None and desc.train:
            self.ops.bn_bwd_sync(desc, Z, A, dA, dZ, dgamma, dbeta, accumulate, B, Cn, self.dp.global_rows(B), self.dp.all_reduce_sum)
        else:
            self.ops.bn_bwd(desc, Z, A, dA, dZ, dgamma, dbeta, accumulate, B, Cn)

    # ---------------------------------------------------------------- grouped Linear building blocks
    # W[g] is [N, K] row-major at arena element offset w_off + g*N*K; bias [N] at b_off + g*N.
    def lin_fwd(self, X: Mat, K, w_off, N, b_off, Y: Mat, M, *, G=1, x_gs=0, relu=False, drop=0.0, salt=0):
        """Y[g] = act(X[g] @ W[g]^T + b[g])   (X[g] = columns g*x_gs.. of X, Y[g] = columns g*N.. of Y)"""
        sp = self.seed_ptr if drop > 0 else None
        if X.is_bf16:
            ybf = Y.is_bf16
            self.ops.gemm_tc(A=X.ptr, lda=X.ld, a_rows=M, a_cols=(G - 1) * x_gs + K, a_mn=0,
                             Bt=self.Wb.data_ptr() + 2 * w_off, ldb=K, b_rows=G * N, b_cols=K, b_mn=0,
                             M=M, N=N, K=K, G=G, a_gk=x_gs, b_gn=N, bias=self.W.data_ptr() + 4 * b_off, bias_gs=N,
                             n_main=N if ybf else 0, out_main=Y.ptr if ybf else None, ld_main=Y.ld, main_gn=N,
                             out_aux=None if ybf else Y.ptr, ld_aux=Y.ld, aux_gn=N, act=1 if relu else 0,
                             drop_p=drop, seed_ptr=sp, salt=salt)
        else:
            self.ops.gemm_f32(A=X.ptr, a_rs=X.ld, a_cs=1, Bt=self.W.data_ptr() + 4 * w_off, b_rs=K, b_cs=1, Cm=Y.ptr, c_rs=Y.ld,
                              M=M, N=N, K=K, G=G, a_gs=x_gs, b_gs=N * K, c_gs=N, bias=self.W.data_ptr() + 4 * b_off, bias_gs=N,
                              act=1 if relu else 0, drop_p=drop, seed_ptr=sp, salt=salt)

    def lin_bwd_w(self, dY: Mat, X: Mat, K, w_off, N, M, *, G=1, x_gs=0):
        """dW[g][n, k] = sum_m dY[g][m, n] * X[g][m, k]  -> gradient arena at w_off (+ g*N*K)."""
        if M == 0:                                               # a replica without rows for this group (routed STAR tower): zero gradient
            self.G[w_off:w_off + G * N * K].zero_()
            return
        gaddr = self.G.data_ptr() + 4 * w_off
        if X.is_bf16:
            assert dY.is_bf16
            self.ops.gemm_tc(A=dY.ptr, lda=dY.ld, a_rows=M, a_cols=G * N, a_mn=1,
                             Bt=X.ptr, ldb=X.ld, b_rows=M, b_cols=(G - 1) * x_gs + K, b_mn=1,
                             M=N, N=K, K=M, G=G, a_gm=N, b_gn=x_gs, n_main=0, out_aux=gaddr, ld_aux=K, aux_gn=N * K,
                             split_k="auto")
        else:
            split = Ops.pick_split(N, K, G, M)
            self.ops.gemm_f32(A=dY.ptr, a_rs=1, a_cs=dY.ld, Bt=X.ptr, b_rs=1, b_cs=X.ld, Cm=gaddr, c_rs=K, M=N, N=K, K=M, G=G,
                              a_gs=N, b_gs=x_gs, c_gs=N * K, split_k=split)

    def lin_bwd_x(self, dY: Mat, K, w_off, N, dX: Mat, M, *, G=1, dx_gs=0, mask: Mat | None = None, mask_gs=0,
                  mask_scale=1.0, accumulate=False):
        """dX[g][m, k] (+)= sum_n dY[g][m, n] * W[g][n, k]; optional ReLU/dropout mask from the forward activation."""
        if dY.is_bf16:
            xbf = dX.is_bf16
            assert mask is None or xbf
            self.ops.gemm_tc(A=dY.ptr, lda=dY.ld, a_rows=M, a_cols=G * N, a_mn=0,
                             Bt=self.Wb.data_ptr() + 2 * w_off, ldb=K, b_rows=G * N, b_cols=K, b_mn=1,
                             M=M, N=K, K=N, G=G, a_gk=N, b_gk=N, n_main=K if xbf else 0,
                             out_main=dX.ptr if xbf else None, ld_main=dX.ld, main_gn=dx_gs,
                             out_aux=None if xbf else dX.ptr, ld_aux=dX.ld, aux_gn=dx_gs,
                             mask=mask.ptr if mask is not None else None, ld_mask=mask.ld if mask is not None else 0,
                             mask_gn=mask_gs, mask_scale=mask_scale, accumulate=1 if accumulate else 0)
        else:
            self.ops.gemm_f32(A=dY.ptr, a_rs=dY.ld, a_cs=1, Bt=self.W.data_ptr() + 4 * w_off, b_rs=1, b_cs=K, Cm=dX.ptr, c_rs=dX.ld,
                              M=M, N=K, K=N, G=G, a_gs=N, b_gs=N * K, c_gs=dx_gs,
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

    def __init__(self, rt: Runtime, tag, G, in_dim, dims, names, *, bn, out_layer, in_groups, g0=0):
        """g0: the arena blocks named by `names` hold MORE groups than this object runs; it runs groups g0 .. g0+G-1 of them
        (STAR runs one tower at a time on that tower's rows, star.py:78-103)."""
        self.rt, self.tag, self.G, self.in_dim, self.dims = rt, tag, G, in_dim, tuple(dims)
        self.names, self.bn, self.out_layer, self.in_groups, self.g0 = names, bn, out_layer, in_groups, g0
        self.salts = [rt.next_salt() for _ in dims]
        # tail0 > 0: the layer-0 pre-activation gradient buffer carries `tail0` extra columns that the caller fills (PLE level 0:
        # the gate / wide-linear logit gradients, whose weights sit right after the experts' in the arena), so that ONE weight-
        # gradient GEMM, ONE input-gradient GEMM and ONE column sum serve experts and gates together
        self.tail0 = 0
        # in_groups that are uniform (group i = input block i feeding experts [i*c, (i+1)*c)) run as ONE grouped launch
        ig = in_groups
        self.uniform = (ig is not None and len(ig) > 1 and all(b == i and e1 - e0 == ig[0][2] - ig[0][1] and e0 == i * (ig[0][2] - ig[0][1])
                                                                for i, (b, e0, e1) in enumerate(ig)))
        if rt.bf16 and (in_dim % 8 or any(d % 8 for d in dims)):
            raise ValueError("the bf16 tensor-core path needs every layer width to be a multiple of 8 (TMA alignment)")

    # arena offsets of layer j's blocks, shifted to this object's first group
    def _k(self, j):
        return self.in_dim if j == 0 else self.dims[j - 1]

    def _oW(self, j, extra=0):
        return self.rt.o(self.names["W"][j], self.g0 * self.dims[j] * self._k(j) + extra)

    def _ob(self, j, extra=0):
        return self.rt.o(self.names["b"][j], self.g0 * self.dims[j] + extra)

    def _vec(self, fn, key, j):
        return fn(self.names[key][j], self.g0 * self.dims[j])

    def can_fuse_tail(self):
        return (not self.bn) and len(self.dims) >= 2 and self.in_groups is not None and len(self.in_groups) == 1 and self.g0 == 0

    def dA0(self, ws: Workspace, B) -> Mat:
        """layer-0 pre-activation gradient [B, G*d0 (+ tail0 columns, pitch padded to 8)]"""
        n = self.G * self.dims[0]
        ld = n + ((self.tail0 + 7) // 8 * 8 if self.tail0 else 0)
        return ws.mat(f"{self.tag}.dA0", B, ld, self.rt.act_dtype)

    def _act(self, ws: Workspace, j, B) -> Mat:
        return ws.mat(f"{self.tag}.A{j}", B, self.G * self.dims[j], self.rt.act_dtype)

    def _layer0(self, X: Mat, Y: Mat, B, fused_act, drop):
        rt, d, K = self.rt, self.dims[0], self.in_dim
        if self.in_groups is None:                               # MLP g reads input block g: one grouped launch
            rt.lin_fwd(X, K, self._oW(0), d, self._ob(0), Y, B, G=self.G, x_gs=K, relu=fused_act, drop=drop, salt=self.salts[0])
        elif self.uniform:                                       # c MLPs per input block: grouped over the blocks
            c = self.in_groups[0][2] - self.in_groups[0][1]
            rt.lin_fwd(X, K, self._oW(0), c * d, self._ob(0), Y, B, G=len(self.in_groups), x_gs=K, relu=fused_act, drop=drop,
                       salt=self.salts[0])
        else:
            for (blk, e0, e1) in self.in_groups:
                rt.lin_fwd(X.cols(blk * K), K, self._oW(0, e0 * d * K), (e1 - e0) * d, self._ob(0, e0 * d), Y.cols(e0 * d), B,
                           relu=fused_act, drop=drop, salt=self.salts[0] + 7919 * e0)

    def fwd_layer0_only(self, ws: Workspace, X: Mat, B):
        """The first (concatenated-N) layer alone, with the training epilogue: used by bench.py to time the dominant GEMM."""
        fused_act = not self.bn
        Y = self._act(ws, 0, B) if fused_act else ws.mat(f"{self.tag}.Z0", B, self.G * self.dims[0])
        self._layer0(X, Y, B, fused_act, self.rt.dropout if fused_act else 0.0)

    def fwd(self, ws: Workspace, X: Mat, B, train) -> Mat:
        rt, G = self.rt, self.G
        use_bn = self.bn and rt.batch_rows(B) != 1               # layer.py:202-204
        drop = rt.dropout if train else 0.0
        prev, prev_d = X, self.in_dim
        for j, d in enumerate(self.dims):
            fused_act = not use_bn
            Y = self._act(ws, j, B) if fused_act else ws.mat(f"{self.tag}.Z{j}", B, G * d)
            if j == 0:
                self._layer0(prev, Y, B, fused_act, drop if fused_act else 0.0)
            else:
                rt.lin_fwd(prev, prev_d, self._oW(j), d, self._ob(j), Y, B, G=G, x_gs=prev_d,
                           relu=fused_act, drop=drop if fused_act else 0.0, salt=self.salts[j])
            if use_bn:
                A = self._act(ws, j, B)
                sm = ws.get(f"{self.tag}.bnsave{j}", (2, G * d))
                desc = rt.ops.bn_desc(self._vec(rt.w, "gamma", j), self._vec(rt.w, "beta", j),
                                      self._vec(rt.b, "rmean", j), self._vec(rt.b, "rvar", j),
                                      sm.data_ptr(), sm.data_ptr() + 4 * G * d, train, True,
                                      drop_p=drop, seed_ptr=rt.seed_ptr if drop > 0 else None, salt=self.salts[j])
                rt.bn_fwd(desc, Y, A, B, G * d)
                Y = A
            prev, prev_d = Y, d
        if self.out_layer:
            L = ws.mat(f"{self.tag}.logit", B, G)
            rt.ops.rowdot_fwd(prev, rt.w(self.names["Wout"], self.g0 * prev_d), rt.w(self.names["bout"], self.g0), L, B, G, prev_d)
            return L
        return prev

    def bwd(self, ws: Workspace, X: Mat, dOut: Mat, B, train, dX: Mat | None, accumulate=False, final_dx=False, post_act_grad=False):
        """dOut: gradient w.r.t. the group's output: dlogits [B, G] fp32 (out_layer), else the gradient of the last
        post-activation [B, G*d_last] (bn) or of the last PRE-activation (no bn: the caller applied the ReLU mask).
        post_act_grad: the caller always hands the post-activation gradient (DCN / DCNv2 heads); a BatchNorm group that skips its
        BatchNorm on a single row (layer.py:202-204) then applies the ReLU / dropout mask itself."""
        rt, G, nl = self.rt, self.G, len(self.dims)
        use_bn = self.bn and rt.batch_rows(B) != 1
        drop = rt.dropout if train else 0.0
        keep = 1.0 / (1.0 - drop) if drop > 0 else 1.0
        d_last = self.dims[-1]
        cur = dOut
        if post_act_grad and self.bn and not use_bn and not self.out_layer:
            # one-row batch: BatchNorm is skipped (layer.py:202-204), the ReLU / dropout mask is applied here; the result feeds the
            # GEMMs below, so it is in the activation dtype (bf16 on the tensor-core path)
            cur = ws.mat(f"{self.tag}.dZ1row", B, G * d_last, rt.act_dtype)
            rt.ops.relu_mask(dOut, self._act(ws, nl - 1, B), cur, B, G * d_last, keep)
        if self.out_layer:
            A_last = self._act(ws, nl - 1, B)
            dA = ws.mat(f"{self.tag}.dA{nl - 1}", B, G * d_last)          # fp32
            rt.ops.rowdot_bwd(A_last, rt.w(self.names["Wout"], self.g0 * d_last), dOut, dA, rt.g(self.names["Wout"], self.g0 * d_last),
                              rt.g(self.names["bout"], self.g0), B, G, d_last)
            cur = dA
            if not use_bn:                                       # one-row batch (layer.py:202-204): no BatchNorm backward to apply the mask
                cur = ws.mat(f"{self.tag}.dZ1row", B, G * d_last, rt.act_dtype) if rt.bf16 else dA
                rt.ops.relu_mask(dA, A_last, cur, B, G * d_last, keep)
        for j in reversed(range(nl)):
            d = self.dims[j]
            prev_d = self.in_dim if j == 0 else self.dims[j - 1]
            if use_bn:
                Z = ws.mat(f"{self.tag}.Z{j}", B, G * d)
                A = self._act(ws, j, B)
                sm = ws.get(f"{self.tag}.bnsave{j}", (2, G * d))
                dZ = ws.mat(f"{self.tag}.dZ{j}", B, G * d, rt.act_dtype)
                desc = rt.ops.bn_desc(self._vec(rt.w, "gamma", j), self._vec(rt.w, "beta", j), None, None,
                                      sm.data_ptr(), sm.data_ptr() + 4 * G * d, train, True, drop_p=drop,
                                      seed_ptr=rt.seed_ptr if drop > 0 else None)
                rt.bn_bwd(desc, Z, A, cur, dZ, self._vec(rt.g, "gamma", j), self._vec(rt.g, "beta", j), False, B, G * d)
                cur = dZ
            # cur is now dZ_j  [B, G*d]
            # Layer 0: the input gradient FIRST - it is what the embedding backward waits for (rt.mark_input_grad), and the bias /
            # weight gradients of this layer can then run next to it.
            if j == 0 and self.tail0 and nl >= 2:                # experts + the caller's tail columns in one go
                n_tot = G * d + self.tail0
                if dX is not None:
                    rt.lin_bwd_x(cur, prev_d, self._oW(j), n_tot, dX, B, accumulate=accumulate)
                    if final_dx:                                 # the caller adds nothing to dX after this call
                        rt.mark_input_grad()
                if X.is_bf16 and cur.is_bf16 and X.ld >= prev_d + 8 and getattr(self, "x_has_ones", False):
                    # X[:, prev_d] == 1: weight and bias gradient from one GEMM
                    rt.ops.wgrad_with_bias_tc(cur, X, prev_d, n_tot, B, rt.G.data_ptr() + 4 * self._oW(j), rt.g(self.names["b"][j]))
                else:
                    rt.ops.colsum(cur, B, n_tot, rt.g(self.names["b"][j]))
                    rt.lin_bwd_w(cur, X, prev_d, self._oW(j), n_tot, B)
                continue
            if j == 0 and self.in_groups is None:
                if dX is not None:
                    rt.lin_bwd_x(cur, prev_d, self._oW(j), d, dX, B, G=G, dx_gs=prev_d, accumulate=accumulate)
                    if final_dx:
                        rt.mark_input_grad()
                rt.ops.colsum(cur, B, G * d, rt.g(self.names["b"][j], self.g0 * d))
                rt.lin_bwd_w(cur, X, prev_d, self._oW(j), d, B, G=G, x_gs=prev_d)
                continue
            rt.ops.colsum(cur, B, G * d, rt.g(self.names["b"][j], self.g0 * d))
            if j == 0 and self.uniform:
                c, ng = self.in_groups[0][2] - self.in_groups[0][1], len(self.in_groups)
                rt.lin_bwd_w(cur, X, prev_d, self._oW(j), c * d, B, G=ng, x_gs=prev_d)
                if dX is not None:
                    rt.lin_bwd_x(cur, prev_d, self._oW(j), c * d, dX, B, G=ng, dx_gs=prev_d, accumulate=accumulate)
            elif j == 0:
                for (blk, e0, e1) in self.in_groups:
                    rt.lin_bwd_w(cur.cols(e0 * d), X.cols(blk * prev_d), prev_d, self._oW(j, e0 * d * prev_d), (e1 - e0) * d, B)
                if dX is not None:
                    seen = set()
                    for (blk, e0, e1) in self.in_groups:
                        acc = accumulate or (blk in seen)
                        seen.add(blk)
                        rt.lin_bwd_x(cur.cols(e0 * d), prev_d, self._oW(j, e0 * d * prev_d), (e1 - e0) * d, dX.cols(blk * prev_d), B,
                                     accumulate=acc)
            else:
                A_prev = self._act(ws, j - 1, B)
                rt.lin_bwd_w(cur, A_prev, prev_d, self._oW(j), d, B, G=G, x_gs=prev_d)
                # next gradient in the activation dtype (bf16 on the tensor-core path: the BatchNorm backward reads either type and
                # keeps its sums in double) - an fp32 gradient here doubled the two HBM passes of every BatchNorm layer
                dA = self.dA0(ws, B) if j == 1 else ws.mat(f"{self.tag}.dA{j - 1}", B, G * prev_d, rt.act_dtype)
                mask = None if use_bn else A_prev
                rt.lin_bwd_x(cur, prev_d, self._oW(j), d, dA, B, G=G, dx_gs=prev_d, mask=mask, mask_gs=prev_d, mask_scale=keep)
                cur = dA
