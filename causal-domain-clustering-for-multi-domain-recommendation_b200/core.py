"""Host-side runtime shared by every model: matrix views over torch storage (device memory only - torch is used for
allocation, streams and state_dict plumbing, never for arithmetic), the per-batch-size workspace, and typed wrappers
that marshal arguments for the C-ABI of libcdcmdr.so.

Nothing in this file computes: every wrapper ends in exactly one `lib.<entry point>` call.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import BnDesc, GemmBf16, GemmF32, MixDesc, PleChain


def _addr(t: torch.Tensor, off: int = 0) -> int:
    return t.data_ptr() + off * t.element_size()


@dataclass(frozen=True)
class Mat:
    """A row-major [rows, cols] window into a flat tensor: element (r, c) lives at t[off + r*ld + c]."""
    t: torch.Tensor
    off: int
    ld: int

    @property
    def ptr(self) -> int:
        return _addr(self.t, self.off)

    def cols(self, c0: int) -> "Mat":
        return Mat(self.t, self.off + c0, self.ld)

    def rows(self, r0: int) -> "Mat":
        return Mat(self.t, self.off + r0 * self.ld, self.ld)

    @property
    def is_bf16(self) -> bool:
        return self.t.dtype == torch.bfloat16


class Workspace:
    """Named device buffers for one batch size; allocated once, reused by every step (graph-capturable)."""

    def __init__(self, device):
        self.device = device
        self.bufs = {}
        self.marks = set()                                      # one-time initialisations done on these buffers

    def get(self, name, shape, dtype=torch.float32, zero=False) -> torch.Tensor:
        key = name
        t = self.bufs.get(key)
        n = 1
        for s in shape:
            n *= int(s)
        if t is None or t.numel() < n or t.dtype != dtype:
            t = (torch.zeros if zero else torch.empty)(max(n, 1), dtype=dtype, device=self.device)
            self.bufs[key] = t
        return t

    def mat(self, name, rows, cols, dtype=torch.float32, zero=False) -> Mat:
        return Mat(self.get(name, (rows, cols), dtype, zero), 0, cols)

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.bufs.values())


class Ops:
    """Argument marshalling for the C-ABI.  One instance per model (holds the device, the current stream getter
    and the shared scratch buffers)."""

    def __init__(self, device):
        self.lib = _lib.load()
        self.device = torch.device(device)
        self._scratch = {}
        self._retired = []
        self.graph_pinned = False                        # set once a CUDA graph holds scratch addresses (Runtime.pin_ws)
        self.launches = 0

    # ---------------------------------------------------------------- plumbing
    @property
    def stream(self) -> int:
        if self.device.type == "cuda":
            return torch.cuda.current_stream(self.device).cuda_stream
        return 0

    def scratch(self, name, nbytes) -> torch.Tensor:
        t = self._scratch.get(name)
        if t is None or t.numel() < nbytes:
            if t is not None and self.graph_pinned:
                self._retired.append(t)                      # a captured graph may still address the smaller buffer
            t = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)
            self._scratch[name] = t
        return t

    # ---------------------------------------------------------------- step state
    def step_state_new(self) -> torch.Tensor:
        st = torch.zeros(_lib.STEP_STATE_BYTES, dtype=torch.uint8, device=self.device)
        self.lib.step_state_init(st.data_ptr(), 0, self.stream)
        return st

    def step_state_set(self, st, step: int):
        self.lib.step_state_init(st.data_ptr(), int(step), self.stream)

    def step_tick(self, st, lr, betas, eps, wd, base_seed):
        self.lib.step_tick(st.data_ptr(), lr, betas[0], betas[1], eps, wd, int(base_seed) & (2 ** 64 - 1), self.stream)

    @staticmethod
    def seed_ptr(st) -> int:
        return st.data_ptr() + _lib.STEP_STATE_SEED_OFFSET

    # ---------------------------------------------------------------- embedding
    def embed_gather(self, x, offsets, table, out_f32: Mat | None, out_bf16: Mat | None, B, F, E, V, oob=None):
        self.lib.embed_gather_fwd(x.data_ptr(), offsets.data_ptr(), table.data_ptr(),
                                  out_f32.ptr if out_f32 is not None else None,
                                  out_bf16.ptr if out_bf16 is not None else None,
                                  out_bf16.ld if out_bf16 is not None else 0,
                                  B, F, E, V, oob.data_ptr() if oob is not None else None, self.stream)

    def embed_plan(self, x, offsets, B, F, V, E) -> torch.Tensor:
        nbytes = self.lib.embed_plan_bytes(B * F, V, E)
        plan = self.scratch("embed_plan", nbytes)            # one buffer for every batch size (grows to the largest seen)
        self.lib.embed_plan_build(x.data_ptr(), offsets.data_ptr(), B, F, V, E, plan.data_ptr(), plan.numel(), self.stream)
        return plan

    def embed_bwd_dense(self, grad_out: Mat, plan, B, F, E, V, grad_table):
        self.lib.embed_bwd_dense(grad_out.ptr, grad_out.ld, plan.data_ptr(), E, B, F, E, V, grad_table.data_ptr(), self.stream)

    def embed_bwd_adam(self, grad_out: Mat, plan, B, F, E, V, table, m, v, l2, st, reg_sumsq=None, lazy=False):
        """grad_out: fp32, or (dense-exact update only) the bf16 row gradients of the replicas' exchange"""
        if lazy:
            self.lib.embed_bwd_adam_sparse_lazy(grad_out.ptr, grad_out.ld, plan.data_ptr(), E, B, F, E, V, table.data_ptr(),
                                                m.data_ptr(), v.data_ptr(), l2, st.data_ptr(), self.stream)
        else:
            fn = self.lib.embed_bwd_adam_dense_exact_g16 if grad_out.is_bf16 else self.lib.embed_bwd_adam_dense_exact
            fn(grad_out.ptr, grad_out.ld, plan.data_ptr(), E, B, F, E, V, table.data_ptr(), m.data_ptr(), v.data_ptr(), l2, st.data_ptr(),
               reg_sumsq.data_ptr() if reg_sumsq is not None else None, self.stream)

    # ---------------------------------------------------------------- fp32 GEMM
    def gemm_f32(self, *, A, a_rs, a_cs, Bt, b_rs, b_cs, Cm, c_rs, M, N, K, G=1, a_gs=0, b_gs=0, c_gs=0,
                 bias=None, bias_gs=0, act=0, mask=None, mask_rs=0, mask_gs=0, mask_scale=1.0,
                 drop_p=0.0, seed_ptr=None, salt=0, accumulate=0, split_k=1):
        """A, Bt, Cm, bias, mask are raw addresses (ints)."""
        if M <= 0 or N <= 0 or G <= 0:
            return
        wsp = None
        if split_k > 1:
            wsp = self.scratch("splitk", 4 * split_k * G * M * N).data_ptr()
        d = GemmF32(A, Bt, Cm, M, N, K, a_rs, a_cs, b_rs, b_cs, c_rs, G, a_gs, b_gs, c_gs, bias, bias_gs, act,
                    mask, mask_rs, mask_gs, mask_scale, drop_p, seed_ptr if drop_p > 0 else None, salt, accumulate,
                    split_k, wsp)
        self.lib.gemm_f32(C.byref(d), self.stream)

    @staticmethod
    def pick_split(M, N, G, K, target_ctas=592, min_k=256):
        ctas = ((M + 63) // 64) * ((N + 63) // 64) * G
        if ctas >= target_ctas or K < 2 * min_k:
            return 1
        # up to 64 slices; a reduction over the B*F token rows of the attention block (K >= 2^18) feeds tiles as small as 64 x 16,
        # where 64 CTAs leave most of the 148 SMs idle: up to 512 slices there
        cap = 512 if K >= (1 << 18) else 64
        s = min((target_ctas + ctas - 1) // ctas, K // min_k, cap)
        return max(1, int(s))

    # ---------------------------------------------------------------- gate mix
    def mix_desc(self, n_gates, n_experts, h, max_sel, desc_t: torch.Tensor, n_pairs=0) -> MixDesc:
        """desc_t: int32 device tensor laid out [gate_col(n_gates) | gate_n(n_gates) | gate_sel(n_gates*max_sel)];
        n_pairs = sum(gate_n), known to the caller on the host."""
        base = desc_t.data_ptr()
        return MixDesc(n_gates, n_experts, h, max_sel, base, base + 4 * n_gates, base + 8 * n_gates, int(n_pairs))

    def gate_mix_fwd(self, d: MixDesc, H: Mat, logits: Mat, out: Mat, probs: torch.Tensor, B):
        self.lib.gate_mix_fwd(C.byref(d), H.ptr, H.ld, logits.ptr, logits.ld, out.ptr, out.ld, probs.data_ptr(), B,
                              1 if H.is_bf16 else 0, self.stream)

    def gate_mix_bwd(self, d: MixDesc, H: Mat, probs, dOut: Mat, dH: Mat, relu_scale, dlogits: Mat, B):
        self.lib.gate_mix_bwd(C.byref(d), H.ptr, H.ld, probs.data_ptr(), dOut.ptr, dOut.ld, dH.ptr, dH.ld, relu_scale,
                              dlogits.ptr, dlogits.ld, B, 1 if H.is_bf16 else 0, self.stream)

    # ---------------------------------------------------------------- batch norm
    def bn_desc(self, gamma, beta, rmean, rvar, save_mean, save_invstd, train, relu, gamma2=None, beta2=None,
                drop_p=0.0, seed_ptr=None, salt=0) -> BnDesc:
        return BnDesc(gamma, beta, gamma2, beta2, rmean, rvar, save_mean, save_invstd, 1 if train else 0,
                      1 if relu else 0, drop_p, seed_ptr if drop_p > 0 else None, salt)

    def bn_fwd(self, d: BnDesc, Z: Mat, A: Mat, B, Cn):
        sc = self.scratch("bn", self.lib.bn_scratch_bytes(Cn))
        self.lib.bn_fwd(C.byref(d), Z.ptr, Z.ld, A.ptr, A.ld, 1 if A.is_bf16 else 0, B, Cn, sc.data_ptr(), self.stream)

    def bn_bwd(self, d: BnDesc, Z: Mat, A: Mat | None, dA: Mat, dZ: Mat, dgamma, dbeta, accumulate, B, Cn):
        sc = self.scratch("bn", self.lib.bn_scratch_bytes(Cn))
        self.lib.bn_bwd(C.byref(d), Z.ptr, Z.ld, A.ptr if A is not None else None, A.ld if A is not None else 0,
                        1 if (A is not None and A.is_bf16) else 0, dA.ptr, dA.ld, 1 if dA.is_bf16 else 0, dZ.ptr, dZ.ld,
                        1 if dZ.is_bf16 else 0, dgamma, dbeta, 1 if accumulate else 0, B, Cn, sc.data_ptr(), self.stream)

    def bn_fwd_sync(self, d: BnDesc, Z: Mat, A: Mat, B, Cn, n_total, all_reduce):
        """Cross-replica BatchNorm forward: local sums -> all_reduce(sums) -> statistics over n_total rows -> normalise."""
        sc = self.scratch("bn", self.lib.bn_scratch_bytes(Cn))
        sums = self.scratch(f"bn_sums", 16 * Cn).view(torch.float64)[:2 * Cn]
        if d.train:
            self.lib.bn_fwd_stats(Z.ptr, Z.ld, B, Cn, sums.data_ptr(), sc.data_ptr(), self.stream)
            all_reduce(sums)
        self.lib.bn_fwd_apply(C.byref(d), Z.ptr, Z.ld, A.ptr, A.ld, 1 if A.is_bf16 else 0, B, n_total, Cn, sums.data_ptr(), self.stream)

    def bn_bwd_sync(self, d: BnDesc, Z: Mat, A: Mat | None, dA: Mat, dZ: Mat, dgamma, dbeta, accumulate, B, Cn, n_total, all_reduce):
        sc = self.scratch("bn", self.lib.bn_scratch_bytes(Cn))
        sums = self.scratch(f"bn_sums", 16 * Cn).view(torch.float64)[:2 * Cn]
        a_ptr, a_ld, a_bf = (A.ptr, A.ld, 1 if A.is_bf16 else 0) if A is not None else (None, 0, 0)
        self.lib.bn_bwd_stats(C.byref(d), Z.ptr, Z.ld, a_ptr, a_ld, a_bf, dA.ptr, dA.ld, 1 if dA.is_bf16 else 0, dgamma, dbeta,
                              1 if accumulate else 0, B, Cn, sums.data_ptr(), sc.data_ptr(), self.stream)
        if d.train:
            all_reduce(sums)
        self.lib.bn_bwd_apply(C.byref(d), Z.ptr, Z.ld, a_ptr, a_ld, a_bf, dA.ptr, dA.ld, 1 if dA.is_bf16 else 0, dZ.ptr, dZ.ld,
                              1 if dZ.is_bf16 else 0, B, n_total, Cn, sums.data_ptr(), sc.data_ptr(), self.stream)

    def copy2d(self, src_addr, lds, dst_addr, ldd, rows, cols, elt_bytes):
        """strided 2-D copy (cdcmdr_permute_rows with the identity permutation)"""
        self.lib.permute_rows(src_addr, lds, None, rows, cols, elt_bytes, dst_addr, ldd, 0, self.stream)

    def copy2d_batched(self, src_addr, src_bs, lds, dst_addr, dst_bs, ldd, batches, rows, cols, elt_bytes):
        """`batches` strided 2-D copies in one launch (element strides)"""
        self.lib.copy2d_batched(src_addr, src_bs, lds, dst_addr, dst_bs, ldd, batches, rows, cols, elt_bytes, self.stream)

    # ---------------------------------------------------------------- Linear(d, 1) heads
    def rowdot_fwd(self, A: Mat, w_addr, b_addr, out: Mat, B, G, d):
        self.lib.rowdot_fwd(A.ptr, A.ld, 1 if A.is_bf16 else 0, w_addr, b_addr, out.ptr, out.ld, B, G, d, self.stream)

    def rowdot_bwd(self, A: Mat, w_addr, dlogit: Mat, dA: Mat | None, dW_addr, db_addr, B, G, d):
        sc = self.scratch("colsum", self.lib.colsum_scratch_bytes(G * d))
        self.lib.rowdot_bwd(A.ptr, A.ld, 1 if A.is_bf16 else 0, w_addr, dlogit.ptr, dlogit.ld, dA.ptr if dA is not None else None,
                            dA.ld if dA is not None else 0, dW_addr, db_addr, B, G, d, sc.data_ptr(), self.stream)

    def cast_f32_bf16(self, src: Mat, dst: Mat, rows, cols):
        self.lib.cast_f32_bf16(src.ptr, src.ld, dst.ptr, dst.ld, rows, cols, self.stream)

    # ---------------------------------------------------------------- bf16 tensor-core GEMM
    def gemm_tc(self, *, A, lda, a_rows, a_cols, a_mn, Bt, ldb, b_rows, b_cols, b_mn, M, N, K, G=1, a_gm=0, a_gk=0, b_gn=0, b_gk=0,
                bias=None, bias_gs=0, n_main=0, out_main=None, ld_main=0, main_gn=0, out_aux=None, ld_aux=0, aux_gn=0, act=0,
                mask=None, ld_mask=0, mask_gn=0, mask_scale=1.0, drop_p=0.0, seed_ptr=None, salt=0, accumulate=0, split_k=1,
                cross_x0=None, cross_x=None, ld_cross=0):
        """A, Bt, outputs, bias, mask are raw addresses.  cross_*: the CrossNetV2 epilogue (cdcmdr.h).  split_k='auto' (weight gradients, n_main == 0): enough K slices to
        fill the machine, partials reduced into out_aux in a fixed order."""
        if M <= 0 or N <= 0 or G <= 0:
            return
        if G > 1 and K % 64 != 0 and (a_gk or b_gk):
            # a 64-wide K block would straddle two groups: one launch per group over that group's slice of the stored matrices
            for g in range(G):
                ao = (g * a_gk * lda if a_mn else g * a_gk) + (g * a_gm if a_mn else g * a_gm * lda)
                bo = (g * b_gk * ldb if b_mn else g * b_gk) + (g * b_gn if b_mn else g * b_gn * ldb)
                self.gemm_tc(A=A + 2 * ao, lda=lda, a_rows=K if a_mn else M, a_cols=M if a_mn else K, a_mn=a_mn,
                             Bt=Bt + 2 * bo, ldb=ldb, b_rows=K if b_mn else N, b_cols=N if b_mn else K, b_mn=b_mn,
                             M=M, N=N, K=K, G=1, bias=bias + 4 * g * bias_gs if bias else None, bias_gs=0, n_main=n_main,
                             out_main=out_main + 2 * g * main_gn if out_main else None, ld_main=ld_main,
                             out_aux=out_aux + 4 * g * aux_gn if out_aux else None, ld_aux=ld_aux, act=act,
                             mask=mask + 2 * g * mask_gn if mask else None, ld_mask=ld_mask, mask_scale=mask_scale,
                             drop_p=drop_p, seed_ptr=seed_ptr, salt=salt + 104729 * g, accumulate=accumulate, split_k=split_k)
            return
        split, part, stride = 1, None, 0
        if split_k == "auto":
            # one persistent CTA per SM walks tiles*split work items: pick the slice count whose (waves x k-blocks per slice) is
            # smallest - e.g. 40 tiles: 8 slices are 3 waves of 128 k-blocks, 11 slices 3 waves of 94
            tiles = ((M + 127) // 128) * ((N + 255) // 256) * G
            num_kb = (K + 63) // 64
            best, want = None, 1
            # reductions over the B*F token rows of the attention block (>= 2^18 rows) feed one or two output tiles: up to one
            # slice per SM there, 48 otherwise
            for cand in range(1, min(148 if num_kb >= 4096 else 48, num_kb) + 1):
                per = (num_kb + cand - 1) // cand
                cost = ((tiles * cand + 147) // 148) * (per + 2)          # + epilogue / pipeline fill per work item
                if best is None or cost < best:
                    best, want = cost, cand
            split = self.lib.gemm_bf16_tc_splits(K, want)
        elif split_k > 1:
            split = self.lib.gemm_bf16_tc_splits(K, split_k)
        dst = out_aux
        if split > 1:
            if n_main != 0 or aux_gn not in (0, M * N) or ld_aux != N:
                raise ValueError("split-K needs a dense fp32 [G, M, N] destination")
            stride = G * M * N
            part = self.scratch("tc_splitk", 4 * split * stride).data_ptr()
            dst = part
        d = GemmBf16(A, lda, a_rows, a_cols, Bt, ldb, b_rows, b_cols, M, N, K, G, a_gm, a_gk, b_gn, b_gk, 1 if a_mn else 0,
                     1 if b_mn else 0, bias, bias_gs, n_main, out_main, ld_main, main_gn, dst, ld_aux, aux_gn, act, mask, ld_mask,
                     mask_gn, mask_scale, drop_p, seed_ptr if drop_p > 0 else None, salt, accumulate if split == 1 else 0,
                     split, stride, 0, cross_x0, cross_x, ld_cross)
        self.lib.gemm_bf16_tc(C.byref(d), self.stream)
        if split > 1:
            self.lib.splitk_reduce(part, stride, split, out_aux, 1, stride, stride, stride, 1 if accumulate else 0, self.stream)

    def ple_chain_ok(self, K0, d0, d1, n_g) -> bool:
        return bool(self.lib.ple_chain_ok(int(K0), int(d0), int(d1), int(n_g)))

    def ple_chain_fwd(self, X: Mat, B, K0, W0_addr, b0_addr, W1_addr, b1_addr, nE, d0, d1, n_g, A0: Mat | None, H: Mat, Lg: Mat | None,
                      drop_p=0.0, seed_ptr=None, salt0=0, salt1=0):
        """First CGC level of PLE in one launch (cdcmdr_ple_chain_fwd): expert layers 0 -> 1 chained through shared memory plus the
        gate / wide-linear logits.  A0 None: inference, the [B, nE*d0] activation is never written."""
        d = PleChain(X.ptr, X.ld, B, K0, W0_addr, b0_addr, W1_addr, b1_addr, nE, d0, d1, n_g,
                     A0.ptr if A0 is not None else None, A0.ld if A0 is not None else 0, H.ptr, H.ld,
                     Lg.ptr if Lg is not None else None, Lg.ld if Lg is not None else 0,
                     drop_p, seed_ptr if drop_p > 0 else None, salt0, salt1)
        self.lib.ple_chain_fwd(C.byref(d), self.stream)

    def wgrad_with_bias_tc(self, dY: Mat, X: Mat, K, M, B, gW_addr, gb_addr):
        """dW[m, k] = sum_b dY[b, m] * X[b, k] (-> gW_addr, [M, K] fp32) and db[m] = sum_b dY[b, m] (-> gb_addr) from ONE product:
        X carries a column of ones at index K (layer.py::_x_mat), so dY^T [X | 1] is [M, K+1] and its last column is the bias
        gradient - the separate column-sum pass over dY (a full HBM read of the largest activation gradient) disappears.
        Split-K partials [split, M, K+1] are reduced straight into the two destinations (fixed order)."""
        N = K + 1
        tiles = ((M + 127) // 128) * ((N + 255) // 256)
        num_kb = (B + 63) // 64
        best, want = None, 1
        for cand in range(1, min(48, num_kb) + 1):
            per = (num_kb + cand - 1) // cand
            cost = ((tiles * cand + 147) // 148) * (per + 2)
            if best is None or cost < best:
                best, want = cost, cand
        split = self.lib.gemm_bf16_tc_splits(B, want)
        ldp = (N + 3) & ~3                                     # partial rows 16-byte aligned: the reduce reads them 128 bits at a time
        stride = M * ldp
        part = self.scratch("tc_splitk", 4 * max(split, 1) * stride).data_ptr()
        d = GemmBf16(dY.ptr, dY.ld, B, M, X.ptr, X.ld, B, N, M, N, B, 1, 0, 0, 0, 0, 1, 1, None, 0, 0, None, 0, 0, part, ldp, 0, 0,
                     None, 0, 0, 1.0, 0.0, None, 0, 0, split, stride, 0, None, None, 0)
        self.lib.gemm_bf16_tc(C.byref(d), self.stream)
        self.lib.splitk_reduce(part, stride, split, gW_addr, M, K, ldp, K, 0, self.stream)
        self.lib.splitk_reduce(part + 4 * K, stride, split, gb_addr, M, 1, ldp, 1, 0, self.stream)

    # ---------------------------------------------------------------- loss
    def sigmoid_select_bce(self, logits, lin: Mat | None, B, T, mode, sel, col, target, pred, psel, loss_sum, dlogits,
                           dlin: Mat | None, inv_batch):
        sc = self.scratch("reduce", self.lib.reduce_scratch_bytes())
        tf32 = 0
        tptr = None
        if target is not None:
            if target.dtype == torch.float32:
                tf32 = 1
            elif target.dtype != torch.int16:
                raise TypeError("targets must be int16 or float32")
            tptr = target.data_ptr()
        self.lib.sigmoid_select_bce(logits.data_ptr(), lin.ptr if lin is not None else None, lin.ld if lin is not None else 0,
                                    B, T, mode, sel.data_ptr() if sel is not None else None, col, tptr, tf32,
                                    pred.data_ptr(), psel.data_ptr() if psel is not None else None,
                                    loss_sum.data_ptr() if loss_sum is not None else None,
                                    dlogits.data_ptr() if dlogits is not None else None,
                                    dlin.ptr if dlin is not None else None, dlin.ld if dlin is not None else 0,
                                    inv_batch, sc.data_ptr(), self.stream)

    def sigmoid_bwd(self, pred, dpred, dlogits, dlin: Mat | None, B, T):
        self.lib.sigmoid_bwd(pred.data_ptr(), dpred.data_ptr(), dlogits.data_ptr(), dlin.ptr if dlin is not None else None,
                             dlin.ld if dlin is not None else 0, B, T, self.stream)

    # ---------------------------------------------------------------- regulariser / optimiser / reductions
    def reg_l2_sum(self, w, coef, coef_scalar, n, out, scratch="reduce"):
        """scratch: name of the partials buffer - a reduction issued on the side stream must not share it with one that may run
        concurrently on the main stream (the fused step's table and dense regulariser sums)."""
        sc = self.scratch(scratch, self.lib.reduce_scratch_bytes())
        self.lib.reg_l2_sum(w.data_ptr(), coef.data_ptr() if coef is not None else None, coef_scalar, n, out.data_ptr(),
                            sc.data_ptr(), self.stream)

    def reg_l2_grad(self, w, coef, coef_scalar, scale, grad, accumulate, n):
        self.lib.reg_l2_grad(w.data_ptr(), coef.data_ptr() if coef is not None else None, coef_scalar, scale, grad.data_ptr(),
                             1 if accumulate else 0, n, self.stream)

    def relu_mask(self, dA: Mat, A: Mat, out: Mat, rows, cols, scale):
        if dA.is_bf16 or A.is_bf16 or out.is_bf16:
            self.lib.relu_mask(dA.ptr, dA.ld, 1 if dA.is_bf16 else 0, A.ptr, A.ld, 1 if A.is_bf16 else 0, out.ptr, out.ld,
                               1 if out.is_bf16 else 0, rows, cols, scale, self.stream)
            return
        self.lib.relu_mask_f32(dA.ptr, dA.ld, A.ptr, A.ld, out.ptr, out.ld, rows, cols, scale, self.stream)

    def adam_dense(self, w, grad, m, v, l2coef, present, n, st):
        self.lib.adam_dense(w.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(),
                            l2coef.data_ptr() if l2coef is not None else None,
                            present.data_ptr() if present is not None else None, n, st.data_ptr(), self.stream)

    def colsum(self, X: Mat, B, Cn, out_addr, accumulate=0):
        sc = self.scratch("colsum", self.lib.colsum_scratch_bytes(Cn))
        self.lib.colsum(X.ptr, X.ld, 1 if X.is_bf16 else 0, B, Cn, out_addr, accumulate, sc.data_ptr(), self.stream)

    # ---------------------------------------------------------------- cross networks / small elementwise stages (fp32)
    def cross_fuse_fwd(self, x0: Mat, x: Mat, xw: Mat, xw_cols, b_addr, out: Mat, B, D):
        """out = x0 * xw + b + x over dense [B, D] blocks (every Mat here has ld == D; xw has ld == xw_cols)"""
        self.lib.cross_fuse_fwd(x0.ptr, x.ptr, xw.ptr, xw_cols, b_addr, out.ptr, B, D, self.stream)

    def cross_fuse_bwd(self, x0: Mat, xw: Mat, xw_cols, dout: Mat, dx0_acc: Mat, dxw: Mat, B, D):
        self.lib.cross_fuse_bwd(x0.ptr, xw.ptr, xw_cols, dout.ptr, dx0_acc.ptr, dxw.ptr, B, D, self.stream)

    def crossmix_combine_fwd(self, x0: Mat, x: Mat, u, g, bias_addr, out: Mat, B, D, n_exp):
        self.lib.crossmix_combine_fwd(x0.ptr, x.ptr, u.data_ptr(), g.data_ptr(), bias_addr, out.ptr, B, D, n_exp, self.stream)

    def crossmix_combine_bwd(self, x0: Mat, u, g, bias_addr, dout: Mat, du, dgate, dx0_acc: Mat, B, D, n_exp):
        self.lib.crossmix_combine_bwd(x0.ptr, u.data_ptr(), g.data_ptr(), bias_addr, dout.ptr, du.data_ptr(), dgate.data_ptr(),
                                      dx0_acc.ptr, B, D, n_exp, self.stream)

    def tanh_fwd(self, t, n):
        self.lib.tanh_fwd(t.data_ptr(), n, self.stream)

    def tanh_bwd(self, y, dy, n):
        self.lib.tanh_bwd(y.data_ptr(), dy.data_ptr(), n, self.stream)

    def softmax_rows_fwd(self, z, ldz, p, ldp, B, n):
        self.lib.softmax_rows_fwd(z.data_ptr(), ldz, p.data_ptr(), ldp, B, n, self.stream)

    def softmax_rows_bwd(self, p, ldp, dp, lddp, dz, lddz, B, n):
        self.lib.softmax_rows_bwd(p.data_ptr(), ldp, dp.data_ptr(), lddp, dz.data_ptr(), lddz, B, n, self.stream)

    def cast_bf16_f32(self, src: Mat, dst: Mat, rows, cols, accumulate=False):
        self.lib.cast_bf16_f32(src.ptr, src.ld, dst.ptr, dst.ld, rows, cols, 1 if accumulate else 0, self.stream)

    def ewise(self, a, b, out, n, op):
        self.lib.ewise_f32(a, b, out, n, op, self.stream)

    def ewise_group(self, a, b, out, n, G, op):
        self.lib.ewise_group_f32(a, b, out, n, G, op, self.stream)

    def add2d(self, a: Mat, out: Mat, rows, cols, accumulate):
        self.lib.add2d_f32(a.ptr, a.ld, out.ptr, out.ld, rows, cols, 1 if accumulate else 0, self.stream)
