"""Host-memory emulator of the C-ABI in include/cdcmdr.h (TEST INFRASTRUCTURE, not product).

Every entry point of libcdcmdr.so is restated here in numpy over HOST pointers, with the same argument lists as the
ctypes binding (package `_lib.SIGNATURES`).  Two uses, both under tests/ only:
  * `-m "not gpu"` tests install it in place of the CUDA library (`_lib.install(HostABI())`) so that the host-side
    logic of the package (arena layout, launch sequencing, backward wiring, optimizer plumbing) is exercised on CPU
    tensors against the model-level oracle (oracle/cdcmdr_oracle.py) and the golden fixtures;
  * `-m gpu` tests use the same functions as the per-kernel oracle: run one entry point on the GPU and the same call
    here on host copies of its inputs.
The product package never imports this module; with the real library on a CPU tensor the package raises.
Each function cites the header declaration it follows; the arithmetic cites the reference where the header does.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
from numpy.lib.stride_tricks import as_strided

F32 = np.float32
BN_EPS, BN_MOM = 1e-5, 0.1


def _arr(ptr, n, dtype):
    if not ptr:
        return None
    dtype = np.dtype(dtype)
    ct = {np.dtype(np.float32): C.c_float, np.dtype(np.float64): C.c_double, np.dtype(np.int32): C.c_int32,
          np.dtype(np.int64): C.c_int64, np.dtype(np.uint8): C.c_uint8, np.dtype(np.uint16): C.c_uint16,
          np.dtype(np.int16): C.c_int16, np.dtype(np.uint64): C.c_uint64, np.dtype(np.uint32): C.c_uint32}[dtype]
    return np.ctypeslib.as_array(C.cast(C.c_void_p(int(ptr)), C.POINTER(ct)), shape=(max(int(n), 1),))


def _mat(ptr, rows, cols, rs, cs=1, dtype=np.float32):
    """Strided 2-D view: element (r, c) at ptr[r*rs + c*cs]."""
    if not ptr:
        return None
    base = _arr(ptr, 1, dtype)
    sz = np.dtype(dtype).itemsize
    return as_strided(base, shape=(int(rows), int(cols)), strides=(int(rs) * sz, int(cs) * sz), writeable=True)


def bf16_to_f32(u16):
    return (u16.astype(np.uint32) << 16).view(np.float32)


def f32_to_bf16(f):
    u = np.ascontiguousarray(f, dtype=np.float32).view(np.uint32)
    rounded = u + (np.uint32(0x7FFF) + ((u >> 16) & 1))           # round to nearest even (__float2bfloat16_rn)
    return (rounded >> 16).astype(np.uint16)


def _act_mat(ptr, rows, cols, ld, is_bf16):
    return _mat(ptr, rows, cols, ld, 1, np.uint16 if is_bf16 else np.float32)


def _rd(m, is_bf16):
    return bf16_to_f32(m) if is_bf16 else m


def _wr(m, val, is_bf16):
    m[...] = f32_to_bf16(val).reshape(m.shape) if is_bf16 else val.astype(np.float32)


def _obj(ref):
    return ref._obj if hasattr(ref, "_obj") else ref.contents


class HostABI:
    is_host_emulator = True

    def __init__(self):
        self._plans = {}
        self._launches = 0
        self.calls = []

    # ------------------------------------------------------------------ library state
    def version(self):
        return 101

    def launch_count(self):
        return self._launches

    def launch_count_reset(self):
        self._launches = 0

    def last_error(self):
        return b""

    def _state(self, st):
        raw = _arr(st, 48, np.uint8)
        return raw[0:8].view(np.int64), raw[8:16].view(np.uint64), raw[16:40].view(np.float32)

    def step_state_init(self, st, step, s):
        t, seed, f = self._state(st)
        t[0] = step; seed[0] = 0; f[:] = 0; f[5] = 1.0
        return 0

    def step_tick(self, st, lr, b1, b2, eps, wd, base_seed, s):
        """cdcmdr_step_tick: torch.optim.Adam bias corrections (run.py:720-721)."""
        t, seed, f = self._state(st)
        t[0] += 1
        step = int(t[0])
        seed[0] = np.uint64((int(base_seed) * 0x9E3779B97F4A7C15 + step) & (2 ** 64 - 1))   # emulator-only seed
        bc1 = 1.0 - float(b1) ** step
        bc2 = 1.0 - float(b2) ** step
        f[0] = F32(float(lr) / bc1); f[1] = F32(b1); f[2] = F32(b2); f[3] = F32(eps); f[4] = F32(wd)
        f[5] = F32(np.sqrt(bc2))
        return 0

    # ------------------------------------------------------------------ embedding (model/layer.py:140-157)
    def embed_gather_fwd(self, x, offsets, table, out_f32, out_bf16, ld_bf16, B, F, E, V, oob, s):
        xs = _mat(x, B, F, F, 1, np.int32)
        off = _arr(offsets, F, np.int64)
        tab = _mat(table, V, E, E)
        idx = xs.astype(np.int64) + off[None, :]
        ok = (idx >= 0) & (idx < V)
        rows = np.where(ok[..., None], tab[np.clip(idx, 0, V - 1)], F32(0)).reshape(B, F * E)
        if oob and not ok.all():
            _arr(oob, 1, np.int32)[0] = 1
        if out_f32:
            _mat(out_f32, B, F * E, F * E)[...] = rows
        if out_bf16:
            _mat(out_bf16, B, F * E, ld_bf16, 1, np.uint16)[...] = f32_to_bf16(rows).reshape(B, F * E)
        return 0

    def embed_plan_bytes(self, n, V, E):
        return 256

    def embed_plan_build(self, x, offsets, B, F, V, E_max, plan, plan_bytes, s):
        xs = _mat(x, B, F, F, 1, np.int32)
        off = _arr(offsets, F, np.int64)
        idx = (xs.astype(np.int64) + off[None, :]).reshape(-1)
        idx = np.where((idx >= 0) & (idx < V), idx, V)
        self._plans[int(plan)] = idx.copy()
        return 0

    def _segment_sums(self, grad_out, ldg, plan, B, F, E, V, g16=False):
        """Sorted-segment scatter-add, ascending (b, f) order inside a row (autograd of nn.Embedding, layer.py:140)."""
        idx = self._plans[int(plan)]
        if g16:                                                  # bf16 row gradients, widened exactly
            go = bf16_to_f32(np.ascontiguousarray(_mat(grad_out, B, F * E, ldg, 1, np.uint16))).reshape(B * F, E)
        else:
            go = _mat(grad_out, B, F * E, ldg).reshape(B * F, E) if ldg == F * E else \
                np.ascontiguousarray(_mat(grad_out, B, F * E, ldg)).reshape(B * F, E)
        g = np.zeros((V + 1, E), dtype=np.float32)
        order = np.argsort(idx, kind="stable")
        rows = idx[order]
        vals = go[order]
        if rows.size:
            starts = np.flatnonzero(np.concatenate([[True], rows[1:] != rows[:-1]]))
            g[rows[starts]] = np.add.reduceat(vals, starts, axis=0)
        touched = np.zeros(V + 1, dtype=bool)
        touched[rows] = True
        return g[:V], touched[:V]

    def embed_bwd_dense(self, grad_out, ldg, plan, E_max, B, F, E, V, grad_table, s):
        g, _ = self._segment_sums(grad_out, ldg, plan, B, F, E, V)
        _mat(grad_table, V, E, E)[...] = g
        return 0

    def _adam(self, w, g, m, v, st, extra_coef):
        """torch _single_tensor_adam with g += extra_coef*w (2*l2 + weight_decay)."""
        _, _, f = self._state(st)
        lr_t, b1, b2, eps, wd, bc2 = (F32(f[i]) for i in range(6))
        g = g + extra_coef * w
        g = g + wd * w
        m[...] = m + (F32(1) - b1) * (g - m)
        v[...] = b2 * v + (F32(1) - b2) * g * g
        denom = np.sqrt(v) / bc2 + eps
        w[...] = w - lr_t * (m / denom)

    def embed_bwd_adam_dense_exact(self, grad_out, ldg, plan, E_max, B, F, E, V, table, m, v, l2, st, reg_sumsq, s):
        g, _ = self._segment_sums(grad_out, ldg, plan, B, F, E, V)
        tab, mm, vv = _mat(table, V, E, E), _mat(m, V, E, E), _mat(v, V, E, E)
        if reg_sumsq:
            _arr(reg_sumsq, 1, np.float64)[0] = np.square(tab.astype(np.float64)).sum()
        self._adam(tab, g, mm, vv, st, F32(2.0) * F32(l2))
        return 0

    def embed_bwd_adam_dense_exact_g16(self, grad_out, ldg, plan, E_max, B, F, E, V, table, m, v, l2, st, reg_sumsq, s):
        g, _ = self._segment_sums(grad_out, ldg, plan, B, F, E, V, g16=True)
        tab, mm, vv = _mat(table, V, E, E), _mat(m, V, E, E), _mat(v, V, E, E)
        if reg_sumsq:
            _arr(reg_sumsq, 1, np.float64)[0] = np.square(tab.astype(np.float64)).sum()
        self._adam(tab, g, mm, vv, st, F32(2.0) * F32(l2))
        return 0

    def embed_bwd_adam_sparse_lazy(self, grad_out, ldg, plan, E_max, B, F, E, V, table, m, v, l2, st, s):
        g, touched = self._segment_sums(grad_out, ldg, plan, B, F, E, V)
        tab, mm, vv = _mat(table, V, E, E), _mat(m, V, E, E), _mat(v, V, E, E)
        w, a, b = tab[touched], mm[touched], vv[touched]
        self._adam(w, g[touched], a, b, st, F32(2.0) * F32(l2))
        tab[touched], mm[touched], vv[touched] = w, a, b
        return 0

    def embed_bwd_adam_sparse_lazy_reg(self, grad_out, ldg, plan, E_max, B, F, E, V, table, m, v, l2, st, reg_running, reg_before, s):
        """lazy update + incrementally maintained sum of squares: *reg_before = running ; running += change of the touched rows"""
        g, touched = self._segment_sums(grad_out, ldg, plan, B, F, E, V)
        tab, mm, vv = _mat(table, V, E, E), _mat(m, V, E, E), _mat(v, V, E, E)
        w, a, b = tab[touched], mm[touched], vv[touched]
        old = np.square(w.astype(np.float64)).sum()
        self._adam(w, g[touched], a, b, st, F32(2.0) * F32(l2))
        tab[touched], mm[touched], vv[touched] = w, a, b
        run = _arr(reg_running, 1, np.float64)
        if reg_before:
            _arr(reg_before, 1, np.float64)[0] = run[0]
        run[0] = run[0] + (np.square(w.astype(np.float64)).sum() - old)
        return 0

    def embed_gather_peer(self, x, offsets, shards, rows_per, out_f32, out_bf16, ld_bf16, B, F, E, V, oob, s):
        """cdcmdr_embed_gather_peer: row r lives in shard r // rows_per (host pointers inside this process)"""
        xs = _mat(x, B, F, F, 1, np.int32)
        off = _arr(offsets, F, np.int64)
        idx = xs.astype(np.int64) + off[None, :]
        ok = (idx >= 0) & (idx < V)
        n_shard = -(-int(V) // int(rows_per))
        ptrs = _arr(shards, n_shard, np.int64)
        full = np.concatenate([_mat(int(ptrs[r]), rows_per, E, E) for r in range(n_shard)], axis=0)
        rows = np.where(ok[..., None], full[np.clip(idx, 0, V - 1)], F32(0)).reshape(B, F * E)
        if oob and not ok.all():
            _arr(oob, 1, np.int32)[0] = 1
        if out_f32:
            _mat(out_f32, B, F * E, F * E)[...] = rows
        if out_bf16:
            _mat(out_bf16, B, F * E, ld_bf16, 1, np.uint16)[...] = f32_to_bf16(rows).reshape(B, F * E)
        return 0

    # ------------------------------------------------------------------ fp32 GEMM (nn.Linear call sites)
    def gemm_f32(self, ref, s):
        p = _obj(ref)
        if p.drop_p > 0:
            raise NotImplementedError("the emulator does not reproduce the dropout hash; test with dropout=0")
        for g in range(p.G):
            A = _mat(p.A + 4 * g * p.a_gs, p.M, p.K, p.a_rs, p.a_cs)
            Bt = _mat(p.Bt + 4 * g * p.b_gs, p.N, p.K, p.b_rs, p.b_cs)
            Cm = _mat(p.C + 4 * g * p.c_gs, p.M, p.N, p.c_rs, 1)
            v = (A.astype(np.float32) @ Bt.astype(np.float32).T).astype(np.float32)
            if p.bias:
                v = v + _arr(p.bias + 4 * g * p.bias_gs, p.N, np.float32)[None, :]
            if p.act == 1:
                v = np.maximum(v, F32(0))
            if p.mask:
                mk = _mat(p.mask + 4 * g * p.mask_gs, p.M, p.N, p.mask_rs, 1)
                v = np.where(mk > 0, v * F32(p.mask_scale), F32(0))
            Cm[...] = (Cm + v) if p.accumulate else v
        return 0

    # ------------------------------------------------------------------ bf16 tensor-core GEMM (same call sites, bf16 operands)
    def gemm_bf16_tc_mode(self, mode):
        return 0

    def gemm_bf16_tc_profile(self, counters):
        return 0

    def gemm_bf16_tc_splits(self, K, want):
        nkb = -(-int(K) // 64)
        s_ = max(1, min(int(want), nkb))
        per = -(-nkb // s_)
        return -(-nkb // per)

    def gemm_bf16_tc(self, ref, s):
        p = _obj(ref)
        if p.drop_p > 0:
            raise NotImplementedError("the emulator does not reproduce the dropout hash; test with dropout=0")
        A = _mat(p.A, p.a_rows, p.a_cols, p.lda, 1, np.uint16)
        Bm = _mat(p.Bt, p.b_rows, p.b_cols, p.ldb, 1, np.uint16)
        split = self.gemm_bf16_tc_splits(p.K, p.split_k) if p.split_k > 1 else 1
        assert split == max(p.split_k, 1), "pass cdcmdr_gemm_bf16_tc_splits(K, want) as split_k"
        nkb = -(-p.K // 64)
        per = -(-nkb // split) * 64
        for g in range(p.G):
            am, ak, bn_, bk = g * p.a_gm, g * p.a_gk, g * p.b_gn, g * p.b_gk
            if p.a_mn_major:
                a = bf16_to_f32(A[ak:ak + p.K, am:am + p.M]).T
            else:
                a = bf16_to_f32(A[am:am + p.M, ak:ak + p.K])
            if p.b_mn_major:
                b = bf16_to_f32(Bm[bk:bk + p.K, bn_:bn_ + p.N]).T
            else:
                b = bf16_to_f32(Bm[bn_:bn_ + p.N, bk:bk + p.K])
            # operands shorter than (M|N, K) because the stored matrix ends: TMA zero-fills
            a = np.pad(a, ((0, p.M - a.shape[0]), (0, p.K - a.shape[1])))
            b = np.pad(b, ((0, p.N - b.shape[0]), (0, p.K - b.shape[1])))
            for z in range(split):
                k0, k1 = z * per, min((z + 1) * per, p.K)
                acc = (a[:, k0:k1].astype(np.float64) @ b[:, k0:k1].astype(np.float64).T).astype(np.float32)
                if getattr(p, "cross_x0", None):
                    # CrossNetV2 epilogue (cdcmdr.h; layer.py:339-343): y = x0 * acc + b + x on bf16 x0 / x; bf16 y, fp32 acc
                    x0 = bf16_to_f32(_mat(p.cross_x0, p.M, p.N, p.ld_cross, 1, np.uint16))
                    xx = bf16_to_f32(_mat(p.cross_x, p.M, p.N, p.ld_cross, 1, np.uint16))
                    bias = _arr(p.bias, p.N, np.float32)[None, :]
                    y = ((x0 * acc + bias) + xx).astype(np.float32)
                    if p.out_aux:
                        _mat(p.out_aux, p.M, p.N, p.ld_aux)[...] = acc
                    _mat(p.out_main, p.M, p.N, p.ld_main, 1, np.uint16)[...] = f32_to_bf16(y).reshape(p.M, p.N)
                    continue
                if p.bias and z == 0:
                    acc = acc + _arr(p.bias + 4 * g * p.bias_gs, p.N, np.float32)[None, :]
                nm = int(p.n_main)
                if nm > 0:
                    v = acc[:, :nm]
                    if p.act == 1:
                        v = np.maximum(v, F32(0))
                    if p.mask:
                        mk = bf16_to_f32(_mat(p.mask + 2 * g * p.mask_gn, p.M, nm, p.ld_mask, 1, np.uint16))
                        v = np.where(mk > 0, v * F32(p.mask_scale), F32(0))
                    om = _mat(p.out_main + 2 * g * p.main_gn, p.M, nm, p.ld_main, 1, np.uint16)
                    if p.accumulate:
                        v = v + bf16_to_f32(om)
                    om[...] = f32_to_bf16(v).reshape(p.M, nm)
                if nm < p.N:
                    oa = _mat(p.out_aux + 4 * (z * p.aux_split_stride + g * p.aux_gn), p.M, p.N - nm, p.ld_aux)
                    v = acc[:, nm:]
                    oa[...] = (oa + v) if (p.accumulate and split == 1) else v
        return 0

    def splitk_reduce(self, part, stride, splits, out, rows, cols, ld_part, ld_out, accumulate, s):
        o = _mat(out, rows, cols, ld_out)
        acc = np.zeros((rows, cols), dtype=np.float32)
        for z in range(splits):
            acc = acc + _mat(part + 4 * z * stride, rows, cols, ld_part)
        o[...] = o + acc if accumulate else acc
        return 0

    def transpose_bf16(self, src, lds, dst, ldd, rows, cols, s):
        _mat(dst, cols, rows, ldd, 1, np.uint16)[...] = _mat(src, rows, cols, lds, 1, np.uint16).T
        return 0

    # ------------------------------------------------------------------ gate softmax + mix (ple.py:106-123, mmoe.py:56-60)
    def _mixdesc(self, ref):
        d = _obj(ref)
        col = _arr(d.gate_col, d.n_gates, np.int32)
        n = _arr(d.gate_n, d.n_gates, np.int32)
        sel = _arr(d.gate_sel, d.n_gates * d.max_sel, np.int32).reshape(d.n_gates, d.max_sel)
        return d, col, n, sel

    def gate_mix_fwd(self, ref, H, ldh, logits, ldl, out, ldo, probs, B, is_bf16, s):
        d, col, n, sel = self._mixdesc(ref)
        h = d.h
        Hm = _rd(_act_mat(H, B, d.n_experts * h, ldh, is_bf16), is_bf16)
        P = _mat(probs, B, d.n_gates * d.max_sel, d.n_gates * d.max_sel)
        P[...] = 0
        for j in range(d.n_gates):
            z = _mat(logits + 4 * int(col[j]), B, int(n[j]), ldl)
            e = np.exp(z - z.max(axis=1, keepdims=True))
            p = (e / e.sum(axis=1, keepdims=True)).astype(np.float32)
            P[:, j * d.max_sel:j * d.max_sel + n[j]] = p
            acc = np.zeros((B, h), dtype=np.float32)
            for k in range(int(n[j])):
                e_ = int(sel[j, k])
                acc += p[:, k:k + 1] * Hm[:, e_ * h:(e_ + 1) * h]
            _wr(_act_mat(out + (2 if is_bf16 else 4) * j * h, B, h, ldo, is_bf16), acc, is_bf16)
        return 0

    def gate_mix_bwd(self, ref, H, ldh, probs, dOut, ldo, dH, lddh, relu_scale, dlogits, lddl, B, is_bf16, s):
        d, col, n, sel = self._mixdesc(ref)
        h = d.h
        Hm = _rd(_act_mat(H, B, d.n_experts * h, ldh, is_bf16), is_bf16)
        dO = _rd(_act_mat(dOut, B, d.n_gates * h, ldo, is_bf16), is_bf16)
        P = _mat(probs, B, d.n_gates * d.max_sel, d.n_gates * d.max_sel)
        acc = np.zeros((B, d.n_experts * h), dtype=np.float32)
        for j in range(d.n_gates):
            nj = int(n[j])
            p = P[:, j * d.max_sel:j * d.max_sel + nj]
            do = dO[:, j * h:(j + 1) * h]
            dp = np.stack([(do * Hm[:, int(sel[j, k]) * h:(int(sel[j, k]) + 1) * h]).sum(1) for k in range(nj)], axis=1)
            dz = p * (dp - (dp * p).sum(1, keepdims=True))
            _mat(dlogits + 4 * int(col[j]), B, nj, lddl)[...] = dz
            for k in range(nj):
                e_ = int(sel[j, k])
                acc[:, e_ * h:(e_ + 1) * h] += p[:, k:k + 1] * do
        if relu_scale > 0:
            acc = np.where(Hm > 0, acc * F32(relu_scale), F32(0))
        _wr(_act_mat(dH, B, d.n_experts * h, lddh, is_bf16), acc, is_bf16)
        return 0

    # ------------------------------------------------------------------ BatchNorm1d (+ReLU) (layer.py:187; star.py:169-181)
    def bn_scratch_bytes(self, Cn):
        return 256

    def bn_fwd(self, ref, Z, ldz, A, lda, a_is_bf16, B, Cn, scratch, s):
        if B == 0 or Cn == 0:
            return 0
        sums = np.zeros(2 * Cn, dtype=np.float64)
        if _obj(ref).train:
            self.bn_fwd_stats(Z, ldz, B, Cn, sums.ctypes.data, scratch, s)
        return self.bn_fwd_apply(ref, Z, ldz, A, lda, a_is_bf16, B, B, Cn, sums.ctypes.data, s)

    def bn_fwd_stats(self, Z, ldz, B, Cn, sums, scratch, s):
        out = _arr(sums, 2 * Cn, np.float64)
        if B == 0:
            out[...] = 0
            return 0
        z64 = _mat(Z, B, Cn, ldz).astype(np.float64)
        out[:Cn] = z64.sum(0)
        out[Cn:] = (z64 * z64).sum(0)
        return 0

    def bn_fwd_apply(self, ref, Z, ldz, A, lda, a_is_bf16, B, n_total, Cn, sums, s):
        p = _obj(ref)
        if p.drop_p > 0:
            raise NotImplementedError("the emulator does not reproduce the dropout hash; test with dropout=0")
        gam = _arr(p.gamma, Cn, np.float32).copy()
        bet = _arr(p.beta, Cn, np.float32).copy()
        if p.gamma2:
            gam = gam * _arr(p.gamma2, Cn, np.float32)
            bet = bet + _arr(p.beta2, Cn, np.float32)
        sm, si = _arr(p.save_mean, Cn, np.float32), _arr(p.save_invstd, Cn, np.float32)
        if p.train:
            sums = _arr(sums, 2 * Cn, np.float64)
            mean = sums[:Cn] / n_total
            var = np.maximum(sums[Cn:] / n_total - mean * mean, 0)
            sm[...] = mean
            si[...] = 1.0 / np.sqrt(var + BN_EPS)
            if p.running_mean:
                rm, rv = _arr(p.running_mean, Cn, np.float32), _arr(p.running_var, Cn, np.float32)
                unb = var * (n_total / (n_total - 1)) if n_total > 1 else var
                rm[...] = F32(1 - BN_MOM) * rm + F32(BN_MOM) * mean.astype(np.float32)
                rv[...] = F32(1 - BN_MOM) * rv + F32(BN_MOM) * unb.astype(np.float32)
        else:
            sm[...] = _arr(p.running_mean, Cn, np.float32)
            si[...] = F32(1.0) / np.sqrt(_arr(p.running_var, Cn, np.float32) + F32(BN_EPS))
        if B == 0:
            return 0
        z = _mat(Z, B, Cn, ldz)
        v = (z - sm) * si * gam + bet
        if p.relu:
            v = np.maximum(v, F32(0))
        _wr(_act_mat(A, B, Cn, lda, a_is_bf16), v.astype(np.float32), a_is_bf16)
        return 0

    def _bn_dy(self, p, A, lda, a_is_bf16, dA, ldda, da_is_bf16, B, Cn):
        dy = _rd(_act_mat(dA, B, Cn, ldda, da_is_bf16), da_is_bf16).astype(np.float32)
        if p.relu:
            a = _rd(_act_mat(A, B, Cn, lda, a_is_bf16), a_is_bf16)
            keep = F32(1.0 / (1.0 - p.drop_p)) if p.drop_p > 0 else F32(1)
            dy = np.where(a > 0, dy * keep, F32(0))
        return dy

    def bn_bwd_stats(self, ref, Z, ldz, A, lda, a_is_bf16, dA, ldda, da_is_bf16, dgamma, dbeta, accumulate, B, Cn, sums, scratch, s):
        p = _obj(ref)
        out = _arr(sums, 2 * Cn, np.float64)
        if B == 0:
            out[...] = 0
            s0 = s1 = np.zeros(Cn, np.float32)
        else:
            z = _mat(Z, B, Cn, ldz)
            dy = self._bn_dy(p, A, lda, a_is_bf16, dA, ldda, da_is_bf16, B, Cn)
            sm, si = _arr(p.save_mean, Cn, np.float32), _arr(p.save_invstd, Cn, np.float32)
            xh = (z - sm) * si
            out[:Cn] = dy.astype(np.float64).sum(0)
            out[Cn:] = (dy.astype(np.float64) * xh).sum(0)
            s0, s1 = out[:Cn].astype(np.float32), out[Cn:].astype(np.float32)
        if dgamma:
            dg = _arr(dgamma, Cn, np.float32)
            dg[...] = dg + s1 if accumulate else s1
        if dbeta:
            db = _arr(dbeta, Cn, np.float32)
            db[...] = db + s0 if accumulate else s0
        return 0

    def bn_bwd_apply(self, ref, Z, ldz, A, lda, a_is_bf16, dA, ldda, da_is_bf16, dZ, lddz, dz_is_bf16, B, n_total, Cn, sums, scratch, s):
        p = _obj(ref)
        if B == 0 or Cn == 0:
            return 0
        z = _mat(Z, B, Cn, ldz)
        dy = self._bn_dy(p, A, lda, a_is_bf16, dA, ldda, da_is_bf16, B, Cn)
        gam = _arr(p.gamma, Cn, np.float32).copy()
        if p.gamma2:
            gam = gam * _arr(p.gamma2, Cn, np.float32)
        sm, si = _arr(p.save_mean, Cn, np.float32), _arr(p.save_invstd, Cn, np.float32)
        sums = _arr(sums, 2 * Cn, np.float64)
        s0, s1 = sums[:Cn].astype(np.float32), sums[Cn:].astype(np.float32)
        if p.train:
            xh = (z - sm) * si
            inv_n = F32(1.0 / n_total)
            dz = gam * si * (dy - s0 * inv_n - xh * s1 * inv_n)
        else:
            dz = dy * gam * si
        _wr(_act_mat(dZ, B, Cn, lddz, dz_is_bf16), dz.astype(np.float32), dz_is_bf16)
        return 0

    def bn_bwd(self, ref, Z, ldz, A, lda, a_is_bf16, dA, ldda, da_is_bf16, dZ, lddz, dz_is_bf16, dgamma, dbeta, accumulate, B, Cn, scratch, s):
        if B == 0 or Cn == 0:
            return 0
        sums = np.zeros(2 * Cn, dtype=np.float64)
        self.bn_bwd_stats(ref, Z, ldz, A, lda, a_is_bf16, dA, ldda, da_is_bf16, dgamma, dbeta, accumulate, B, Cn, sums.ctypes.data, scratch, s)
        return self.bn_bwd_apply(ref, Z, ldz, A, lda, a_is_bf16, dA, ldda, da_is_bf16, dZ, lddz, dz_is_bf16, B, B, Cn, sums.ctypes.data, scratch, s)

    # ------------------------------------------------------------------ Linear(d, 1) heads (layer.py:192-193)
    def rowdot_fwd(self, A, lda, a_is_bf16, w, bias, out, ldo, B, G, d, s):
        a = _rd(_act_mat(A, B, G * d, lda, a_is_bf16), a_is_bf16).reshape(B, G, d)
        ww = _arr(w, G * d, np.float32).reshape(G, d)
        v = np.einsum("bgk,gk->bg", a.astype(np.float32), ww).astype(np.float32)
        if bias:
            v = v + _arr(bias, G, np.float32)[None, :]
        _mat(out, B, G, ldo)[...] = v
        return 0

    def rowdot_bwd(self, A, lda, a_is_bf16, w, dlogit, ldl, dA, ldda, dW, dbias, B, G, d, scratch, s):
        a = _rd(_act_mat(A, B, G * d, lda, a_is_bf16), a_is_bf16).reshape(B, G, d)
        ww = _arr(w, G * d, np.float32).reshape(G, d)
        dl = _mat(dlogit, B, G, ldl)
        if dA:
            _mat(dA, B, G * d, ldda)[...] = (dl[:, :, None] * ww[None]).reshape(B, G * d)
        if dW:
            _arr(dW, G * d, np.float32)[...] = np.einsum("bg,bgk->gk", dl.astype(np.float64), a.astype(np.float64)).reshape(-1)
        if dbias:
            _arr(dbias, G, np.float32)[...] = dl.astype(np.float64).sum(0)
        return 0

    # ------------------------------------------------------------------ sigmoid / selection / BCE (layer.py:48-56; cdc.py:99-111; run.py:723)
    def reduce_scratch_bytes(self):
        return 8192

    def sigmoid_select_bce(self, logits, lin, ld_lin, B, T, mode, sel, col, target, target_is_f32, pred, psel, loss_sum,
                           dlogits, dlin, ld_dlin, inv_batch, scratch, s):
        z = _mat(logits, B, T, T).astype(np.float32)
        if lin:
            z = z + _mat(lin, B, 1, ld_lin)
        y = (F32(1) / (F32(1) + np.exp(-z))).astype(np.float32)
        _mat(pred, B, T, T)[...] = y
        ar = np.arange(B)
        bad = np.zeros(B, dtype=bool)
        if mode == 0:
            c = _arr(sel, B, np.int64).astype(np.int64)
            bad = (c < 0) | (c >= T)                              # out-of-range selection: NaN prediction / loss, no gradient
            c = np.where(bad, 0, c)
            ps = np.where(bad, F32(np.nan), y[ar, c]).astype(np.float32)
        elif mode == 1:
            c = np.full(B, col, dtype=np.int64)
            ps = y[ar, c]
        elif mode == 2:
            ps = (y.sum(1) / F32(T)).astype(np.float32)
        else:
            ps = np.zeros(B, dtype=np.float32)
        if psel:
            _arr(psel, B, np.float32)[...] = ps
        if target:
            tg = _arr(target, B, np.float32 if target_is_f32 else np.int16).astype(np.float32)
            with np.errstate(divide="ignore"):
                lp = np.maximum(np.log(ps), F32(-100))
                l1p = np.maximum(np.log1p(-ps), F32(-100))
            with np.errstate(invalid="ignore"):
                _arr(loss_sum, 1, np.float64)[0] = np.where(bad, np.nan, -(tg * lp + (F32(1) - tg) * l1p)).astype(np.float64).sum()
            if dlogits:
                dps = (ps - tg) / np.maximum((F32(1) - ps) * ps, F32(1e-12)) * F32(inv_batch)
                dz = np.zeros((B, T), dtype=np.float32)
                if mode == 2:
                    dz = (dps / F32(T))[:, None] * y * (F32(1) - y)
                else:
                    ys = y[ar, c]
                    with np.errstate(invalid="ignore"):
                        dz[ar, c] = np.where(bad, F32(0), dps * ys * (F32(1) - ys))
                _mat(dlogits, B, T, T)[...] = dz
                if dlin:
                    _mat(dlin, B, 1, ld_dlin)[:, 0] = dz.sum(1)
        return 0

    def sigmoid_bwd(self, pred, dpred, dlogits, dlin, ld_dlin, B, T, s):
        y = _mat(pred, B, T, T)
        dz = _mat(dpred, B, T, T) * y * (F32(1) - y)
        _mat(dlogits, B, T, T)[...] = dz
        if dlin:
            _mat(dlin, B, 1, ld_dlin)[:, 0] = dz.sum(1)
        return 0

    # ------------------------------------------------------------------ regulariser / Adam / reductions
    def reg_l2_sum(self, w, coef, coef_scalar, n, out, scratch, s):
        ww = _arr(w, n, np.float32).astype(np.float64)
        c = _arr(coef, n, np.float32).astype(np.float64) if coef else float(F32(coef_scalar))
        _arr(out, 1, np.float64)[0] = (c * ww * ww).sum()
        return 0

    def reg_l2_grad(self, w, coef, coef_scalar, scale, grad, accumulate, n, s):
        ww = _arr(w, n, np.float32)
        c = _arr(coef, n, np.float32) if coef else F32(coef_scalar)
        g = _arr(grad, n, np.float32)
        v = F32(scale) * F32(2) * c * ww
        g[...] = g + v if accumulate else v
        return 0

    def relu_mask_f32(self, dA, ldda, A, lda, out, ldo, rows, cols, scale, s):
        _mat(out, rows, cols, ldo)[...] = np.where(_mat(A, rows, cols, lda) > 0, _mat(dA, rows, cols, ldda) * F32(scale), F32(0))
        return 0

    def relu_mask(self, dA, ldda, da_bf16, A, lda, a_bf16, out, ldo, out_bf16, rows, cols, scale, s):
        a = _rd(_act_mat(A, rows, cols, lda, a_bf16), a_bf16)
        d = _rd(_act_mat(dA, rows, cols, ldda, da_bf16), da_bf16)
        _wr(_act_mat(out, rows, cols, ldo, out_bf16), np.where(a > 0, d * F32(scale), F32(0)).astype(np.float32), out_bf16)
        return 0

    def adam_dense(self, w, grad, m, v, l2coef, present, n, st, s):
        ww, g, mm, vv = (_arr(p, n, np.float32) for p in (w, grad, m, v))
        coef = F32(2) * _arr(l2coef, n, np.float32) if l2coef else F32(0)
        if present:
            pr = _arr(present, n, np.uint8) != 0
            w2, m2, v2 = ww[pr], mm[pr], vv[pr]
            self._adam(w2, g[pr], m2, v2, st, coef[pr] if l2coef else coef)
            ww[pr], mm[pr], vv[pr] = w2, m2, v2
        else:
            self._adam(ww, g.copy(), mm, vv, st, coef)
        return 0

    def colsum_scratch_bytes(self, Cn):
        return 256

    def colsum(self, X, ld, is_bf16, B, Cn, out, accumulate, scratch, s):
        x = _rd(_act_mat(X, B, Cn, ld, is_bf16), is_bf16)
        t = x.astype(np.float64).sum(0).astype(np.float32)
        o = _arr(out, Cn, np.float32)
        o[...] = o + t if accumulate else t
        return 0

    def cast_f32_bf16(self, src, lds, dst, ldd, rows, cols, s):
        _mat(dst, rows, cols, ldd, 1, np.uint16)[...] = f32_to_bf16(np.ascontiguousarray(_mat(src, rows, cols, lds))).reshape(rows, cols)
        return 0

    def cast_bf16_f32(self, src, lds, dst, ldd, rows, cols, accumulate, s):
        v = bf16_to_f32(np.ascontiguousarray(_mat(src, rows, cols, lds, 1, np.uint16)))
        d = _mat(dst, rows, cols, ldd)
        d[...] = d + v if accumulate else v
        return 0

    def ewise_group_f32(self, a, b, out, n, G, op, s):
        aa = _arr(a, n * G, np.float32)[:n * G].reshape(G, n)
        if op in (0, 1, 4):
            o = _arr(out, n * G, np.float32)[:n * G].reshape(G, n)
            bb = _arr(b, n, np.float32)[:n] if b else None
            o[...] = aa * bb[None, :] if op == 0 else (aa + bb[None, :] if op == 1 else o + aa)
            return 0
        o = _arr(out, n, np.float32)[:n]
        if op == 2:
            bb = _arr(b, n * G, np.float32)[:n * G].reshape(G, n)
            acc = np.zeros(n, dtype=np.float32)
            for g in range(G):
                acc = acc + aa[g] * bb[g]
            o[...] = acc
        else:
            acc = np.zeros(n, dtype=np.float32)
            for g in range(G):
                acc = acc + aa[g]
            o[...] = o + acc
        return 0

    def ewise_f32(self, a, b, out, n, op, s):
        aa, o = _arr(a, n, np.float32), _arr(out, n, np.float32)
        bb = _arr(b, n, np.float32) if b else None
        if op == 0:
            o[...] = aa * bb
        elif op == 1:
            o[...] = aa + bb
        elif op == 2:
            o[...] = o + aa * bb
        else:
            o[...] = o + aa
        return 0

    def add2d_f32(self, a, lda, out, ldo, rows, cols, accumulate, s):
        o = _mat(out, rows, cols, ldo)
        v = _mat(a, rows, cols, lda)
        o[...] = o + v if accumulate else v
        return 0

    # ------------------------------------------------------------------ cross networks (layer.py:495-515, 332-343, 380-407)
    def cross_fuse_fwd(self, x0, x, xw, xw_cols, b, out, B, D, s):
        w = _mat(xw, B, xw_cols, xw_cols)
        _mat(out, B, D, D)[...] = _mat(x0, B, D, D) * w + _arr(b, D, np.float32)[None, :] + _mat(x, B, D, D)
        return 0

    def cross_fuse_bwd(self, x0, xw, xw_cols, dout, dx0_acc, dxw, B, D, s):
        g = _mat(dout, B, D, D)
        w = _mat(xw, B, xw_cols, xw_cols)
        X0 = _mat(x0, B, D, D)
        if xw_cols == 1 and D != 1:
            _mat(dxw, B, 1, 1)[:, 0] = (g * X0).sum(1)
        else:
            _mat(dxw, B, D, D)[...] = g * X0
        _mat(dx0_acc, B, D, D)[...] += g * w
        return 0

    def crossmix_combine_fwd(self, x0, x, u, g, bias, out, B, D, n_exp, s):
        U = _arr(u, n_exp * B * D, np.float32).reshape(n_exp, B, D)
        G = _mat(g, B, n_exp, n_exp)
        X0 = _mat(x0, B, D, D)
        bb = _arr(bias, D, np.float32)
        acc = np.zeros((B, D), dtype=np.float32)
        for e in range(n_exp):
            acc += G[:, e:e + 1] * (X0 * (U[e] + bb))
        _mat(out, B, D, D)[...] = _mat(x, B, D, D) + acc
        return 0

    def crossmix_combine_bwd(self, x0, u, g, bias, dout, du, dgate, dx0_acc, B, D, n_exp, s):
        U = _arr(u, n_exp * B * D, np.float32).reshape(n_exp, B, D)
        dU = _arr(du, n_exp * B * D, np.float32).reshape(n_exp, B, D)
        G = _mat(g, B, n_exp, n_exp)
        X0 = _mat(x0, B, D, D)
        bb = _arr(bias, D, np.float32)
        go = _mat(dout, B, D, D)
        dG = _mat(dgate, B, n_exp, n_exp)
        acc = _mat(dx0_acc, B, D, D)
        for e in range(n_exp):
            ub = U[e] + bb
            dU[e] = go * G[:, e:e + 1] * X0
            dG[:, e] = (go * X0 * ub).sum(1)
            acc[...] += go * G[:, e:e + 1] * ub
        return 0

    def tanh_fwd(self, x, n, s):
        a = _arr(x, n, np.float32)
        a[...] = np.tanh(a)
        return 0

    def tanh_bwd(self, y, dy, n, s):
        t, d = _arr(y, n, np.float32), _arr(dy, n, np.float32)
        d[...] = d * (F32(1) - t * t)
        return 0

    def softmax_rows_fwd(self, z, ldz, p, ldp, B, n, s):
        zz = _mat(z, B, n, ldz)
        e = np.exp(zz - zz.max(axis=1, keepdims=True))
        _mat(p, B, n, ldp)[...] = e / e.sum(axis=1, keepdims=True)
        return 0

    def softmax_rows_bwd(self, p, ldp, dp, lddp, dz, lddz, B, n, s):
        P, dP = _mat(p, B, n, ldp), _mat(dp, B, n, lddp)
        _mat(dz, B, n, lddz)[...] = P * (dP - (dP * P).sum(1, keepdims=True))
        return 0

    # ------------------------------------------------------------------ routing (star.py:84-86,107,112-114; cdc.py:105)
    def route_scratch_bytes(self, B, n_group):
        return 256

    def route_partition(self, group, B, n_group, perm, counts, group_start, scratch, s):
        g = _arr(group, B, np.int64)[:B] if B > 0 else np.zeros(0, np.int64)
        pm = _arr(perm, B, np.int32)
        cnt = _arr(counts, n_group, np.int32)
        gs = _arr(group_start, n_group + 1, np.int32)
        off = 0
        for k in range(n_group):
            rows = np.flatnonzero(g == k)
            gs[k] = off
            cnt[k] = rows.size
            pm[off:off + rows.size] = rows
            off += rows.size
        gs[n_group] = off
        return 0

    def bce_segments_scratch_bytes(self, n_seg):
        return 256

    def bce_segments(self, pred, ld_pred, target, target_is_f32, seg_start, n_seg, out_mean, scratch, s):
        st = _arr(seg_start, n_seg + 1, np.int64)
        n = int(st[n_seg])
        p = _mat(pred, n, 1, ld_pred)[:, 0].astype(np.float32) if n else np.zeros(0, np.float32)
        tg = _arr(target, n, np.float32 if target_is_f32 else np.int16).astype(np.float32) if n else np.zeros(0, np.float32)
        out = _arr(out_mean, n_seg, np.float32)
        with np.errstate(divide="ignore"):
            lp = np.maximum(np.log(p), F32(-100))
            l1p = np.maximum(np.log1p(-p), F32(-100))
        term = (-(tg * lp + (F32(1) - tg) * l1p)).astype(np.float64)
        for k in range(n_seg):
            a, b = int(st[k]), int(st[k + 1])
            out[k] = np.float32(term[a:b].sum() / (b - a)) if b > a else np.float32(np.nan)
        return 0

    def copy2d_batched(self, src, src_bs, lds, dst, dst_bs, ldd, batches, rows, cols, elt_bytes, s):
        dt = {2: np.uint16, 4: np.uint32}[elt_bytes]
        sz = np.dtype(dt).itemsize
        if batches <= 0 or rows <= 0 or cols <= 0:
            return 0
        S = as_strided(_mat(src, 1, 1, lds, 1, dt), (batches, rows, cols), (src_bs * sz, lds * sz, sz))
        Dm = as_strided(_mat(dst, 1, 1, ldd, 1, dt), (batches, rows, cols), (dst_bs * sz, ldd * sz, sz), writeable=True)
        Dm[...] = S
        return 0

    def permute_rows(self, src, lds, perm, n, cols, elt_bytes, dst, ldd, scatter, s):
        dt = {2: np.uint16, 4: np.uint32, 8: np.uint64}[elt_bytes]
        S, Dm = _mat(src, 1, 1, lds, 1, dt), _mat(dst, 1, 1, ldd, 1, dt)
        pm = _arr(perm, n, np.int32)[:n].astype(np.int64) if perm else np.arange(n, dtype=np.int64)
        sz = np.dtype(dt).itemsize
        rows_src = int(pm.max()) + 1 if (n and not scatter) else n
        rows_dst = int(pm.max()) + 1 if (n and scatter) else n
        S = as_strided(S, (rows_src, cols), (lds * sz, sz))
        Dm = as_strided(Dm, (rows_dst, cols), (ldd * sz, sz), writeable=True)
        if scatter:
            Dm[pm] = S[:n]
        else:
            Dm[...] = S[pm]
        return 0

    def domain_to_group(self, x, B, F, domain_idx, d2g, n_domain, groups, s):
        xs = _mat(x, B, F, F, 1, np.int32)
        d = xs[:, domain_idx].astype(np.int64)
        m = _arr(d2g, n_domain, np.int64)
        _arr(groups, B, np.int64)[:B] = np.where((d >= 0) & (d < n_domain), m[np.clip(d, 0, n_domain - 1)], -1)
        return 0

    # ------------------------------------------------------------------ N3: field self-attention (model/layer.py:58-84)
    # The core of torch.nn.MultiheadAttention as F.multi_head_attention_forward computes it (no masks, need_weights irrelevant):
    # q, k, v split per head, softmax(q k^T / sqrt(dh)) over the L tokens of one sample, weighted sum of v.
    def _attn_views(self, ptr, ld, B, L, H, dh):
        A = H * dh
        m = _mat(ptr, B * L, 3 * A, ld).reshape(B, L, 3, H, dh)
        return m[:, :, 0].transpose(0, 2, 1, 3), m[:, :, 1].transpose(0, 2, 1, 3), m[:, :, 2].transpose(0, 2, 1, 3)     # [B, H, L, dh]

    def attn_fwd(self, qkv, ld, out, ldo, probs, B, L, H, dh, scale, drop_p, seed_dev, salt, s):
        if drop_p > 0:
            raise NotImplementedError("the emulator does not reproduce the dropout hash; test with dropout=0")
        if B <= 0:
            return 0
        q, k, v = self._attn_views(qkv, ld, B, L, H, dh)
        sc = np.einsum("bhid,bhjd->bhij", q, k).astype(np.float32) * F32(scale)
        sc = sc - sc.max(axis=-1, keepdims=True)
        e = np.exp(sc).astype(np.float32)
        p = (e / e.sum(axis=-1, keepdims=True, dtype=np.float32)).astype(np.float32)
        if probs:
            _arr(probs, B * H * L * L, np.float32)[:B * H * L * L] = p.reshape(-1)
        o = np.einsum("bhij,bhjd->bhid", p, v).astype(np.float32)                       # [B, H, L, dh]
        _mat(out, B * L, H * dh, ldo)[...] = o.transpose(0, 2, 1, 3).reshape(B * L, H * dh)
        return 0

    def attn_bwd(self, qkv, ld, probs, dout, lddo, dqkv, lddq, B, L, H, dh, scale, drop_p, seed_dev, salt, s):
        if drop_p > 0:
            raise NotImplementedError("the emulator does not reproduce the dropout hash; test with dropout=0")
        if B <= 0:
            return 0
        A = H * dh
        q, k, v = self._attn_views(qkv, ld, B, L, H, dh)
        p = _arr(probs, B * H * L * L, np.float32)[:B * H * L * L].reshape(B, H, L, L)
        do = _mat(dout, B * L, A, lddo).reshape(B, L, H, dh).transpose(0, 2, 1, 3)      # [B, H, L, dh]
        dv = np.einsum("bhij,bhid->bhjd", p, do)
        dp = np.einsum("bhid,bhjd->bhij", do, v)
        ds = (p * (dp - (p * dp).sum(axis=-1, keepdims=True)) * F32(scale)).astype(np.float32)
        dq = np.einsum("bhij,bhjd->bhid", ds, k)
        dk = np.einsum("bhij,bhid->bhjd", ds, q)
        g = _mat(dqkv, B * L, 3 * A, lddq)
        for idx, t in enumerate((dq, dk, dv)):
            g[:, idx * A:(idx + 1) * A] = t.astype(np.float32).transpose(0, 2, 1, 3).reshape(B * L, A)
        return 0

    def attn_pool_fwd(self, z, w, lin, ld_lin, accumulate, B, n, s):
        if B <= 0:
            return 0
        zz = _mat(z, B, n, n)
        v = (np.maximum(zz, 0) @ _arr(w, n, np.float32)[:n]).astype(np.float32)
        o = _mat(lin, B, 1, ld_lin)
        o[:, 0] = (o[:, 0] + v) if accumulate else v
        return 0

    def attn_pool_scratch_bytes(self, B, n):
        return 256

    def attn_pool_bwd(self, z, w, dlin, ld_dlin, dz, dw, B, n, scratch, s):
        if B <= 0:
            _arr(dw, n, np.float32)[:n] = 0
            return 0
        zz = _mat(z, B, n, n)
        ww = _arr(w, n, np.float32)[:n]
        g = _mat(dlin, B, 1, ld_dlin)[:, 0]
        _mat(dz, B, n, n)[...] = np.where(zz > 0, g[:, None] * ww[None, :], F32(0)).astype(np.float32)
        _arr(dw, n, np.float32)[:n] = (g.astype(np.float64)[:, None] * np.maximum(zz, 0).astype(np.float64)).sum(axis=0).astype(np.float32)
        return 0

    # ---- the same stages on bf16 token matrices (cdcmdr_attn_*_bf16): fp32 arithmetic on bf16-rounded inputs, bf16 outputs, no
    # stored probabilities (the backward recomputes the softmax)
    def _attn_views_bf16(self, ptr, ld, B, L, H, dh):
        A = H * dh
        m = bf16_to_f32(_mat(ptr, B * L, 3 * A, ld, 1, np.uint16)).reshape(B, L, 3, H, dh)
        return m[:, :, 0].transpose(0, 2, 1, 3), m[:, :, 1].transpose(0, 2, 1, 3), m[:, :, 2].transpose(0, 2, 1, 3)

    @staticmethod
    def _attn_softmax(q, k, scale):
        sc = np.einsum("bhid,bhjd->bhij", q, k).astype(np.float32) * F32(scale)
        sc = sc - sc.max(axis=-1, keepdims=True)
        e = np.exp(sc).astype(np.float32)
        return (e / e.sum(axis=-1, keepdims=True, dtype=np.float32)).astype(np.float32)

    def attn_fwd_bf16(self, qkv, ld, out, ldo, B, L, H, dh, scale, drop_p, seed_dev, salt, s):
        if drop_p > 0:
            raise NotImplementedError("the emulator does not reproduce the dropout hash; test with dropout=0")
        if B <= 0:
            return 0
        q, k, v = self._attn_views_bf16(qkv, ld, B, L, H, dh)
        p = self._attn_softmax(q, k, scale)
        o = np.einsum("bhij,bhjd->bhid", p, v).astype(np.float32)
        _mat(out, B * L, H * dh, ldo, 1, np.uint16)[...] = f32_to_bf16(o.transpose(0, 2, 1, 3).reshape(B * L, H * dh)).reshape(B * L, H * dh)
        return 0

    def attn_bwd_bf16(self, qkv, ld, dout, lddo, dqkv, lddq, B, L, H, dh, scale, drop_p, seed_dev, salt, s):
        if drop_p > 0:
            raise NotImplementedError("the emulator does not reproduce the dropout hash; test with dropout=0")
        if B <= 0:
            return 0
        A = H * dh
        q, k, v = self._attn_views_bf16(qkv, ld, B, L, H, dh)
        p = self._attn_softmax(q, k, scale)
        do = bf16_to_f32(_mat(dout, B * L, A, lddo, 1, np.uint16)).reshape(B, L, H, dh).transpose(0, 2, 1, 3)
        dv = np.einsum("bhij,bhid->bhjd", p, do)
        dp = np.einsum("bhid,bhjd->bhij", do, v)
        ds = (p * (dp - (p * dp).sum(axis=-1, keepdims=True)) * F32(scale)).astype(np.float32)
        dq = np.einsum("bhij,bhjd->bhid", ds, k)
        dk = np.einsum("bhij,bhid->bhjd", ds, q)
        g = _mat(dqkv, B * L, 3 * A, lddq, 1, np.uint16)
        for idx, t in enumerate((dq, dk, dv)):
            g[:, idx * A:(idx + 1) * A] = f32_to_bf16(t.astype(np.float32).transpose(0, 2, 1, 3).reshape(B * L, A)).reshape(B * L, A)
        return 0

    def attn_pool_fwd_bf16(self, z, w, lin, ld_lin, accumulate, B, n, s):
        if B <= 0:
            return 0
        zz = bf16_to_f32(_mat(z, B, n, n, 1, np.uint16))
        v = (np.maximum(zz, 0) @ _arr(w, n, np.float32)[:n]).astype(np.float32)
        o = _mat(lin, B, 1, ld_lin)
        o[:, 0] = (o[:, 0] + v) if accumulate else v
        return 0

    def attn_pool_bwd_bf16(self, z, w, dlin, ld_dlin, dz, dw, B, n, scratch, s):
        if B <= 0:
            _arr(dw, n, np.float32)[:n] = 0
            return 0
        zz = bf16_to_f32(_mat(z, B, n, n, 1, np.uint16))
        ww = _arr(w, n, np.float32)[:n]
        g = _mat(dlin, B, 1, ld_dlin)[:, 0]
        _mat(dz, B, n, n, 1, np.uint16)[...] = f32_to_bf16(np.where(zz > 0, g[:, None] * ww[None, :], F32(0)).astype(np.float32)).reshape(B, n)
        _arr(dw, n, np.float32)[:n] = (g.astype(np.float64)[:, None] * np.maximum(zz, 0).astype(np.float64)).sum(axis=0).astype(np.float32)
        return 0

    # ---- (e) peer-memory all-reduce: needs NVLink peer mappings between processes; the CPU tests exchange through gloo instead
    def peer_allreduce_bytes(self, world, max_n):
        return 2 * world * max_n * 8 + 2 * world * 8

    def peer_allreduce_f64(self, peer_bufs, rank, world, inp, out, n, max_n, seq, s):
        if world != 1:
            raise NotImplementedError("the host emulator has no peer memory: multi-rank CPU tests use torch.distributed (gloo)")
        _arr(out, n, np.float64)[:n] = _arr(inp, n, np.float64)[:n]
        return 0

    def peer_allreduce_f32_chunk(self, world, n):
        return 0 if (world < 1 or n < 1) else ((-(-n // world)) + 3) & ~3

    def peer_allreduce_f32(self, inbox, outbox, peer_flags, my_inbox, my_outbox, rank, world, inp, out, n, seqs2, s):
        if world != 1:
            raise NotImplementedError("the host emulator has no peer memory: multi-rank CPU tests use torch.distributed (gloo)")
        _arr(out, n, np.float32)[:n] = _arr(inp, n, np.float32)[:n]
        return 0

    # ---- (e) fused embedding exchange: the "peers" are host buffers inside this process (one emulator call per rank, in any order)
    def peer_barrier(self, peer_flags, rank, world, slot, n_slots, seqs, s):
        if world != 1:
            raise NotImplementedError("the host emulator has no peer memory: a barrier between ranks cannot be emulated in one call")
        q = _arr(seqs, n_slots, np.uint64)
        q[slot] += 1
        _arr(_arr(peer_flags, 1, np.int64)[0], n_slots * world, np.uint64)[slot * world + rank] = q[slot]
        return 0

    def dp_push_ids(self, x, B, F, recv_ids, fbound, rank, world, s):
        if B == 0:
            return 0
        xs = _mat(x, B, F, F, 1, np.int32)
        fb = _arr(fbound, world + 1, np.int32)
        ptrs = _arr(recv_ids, world, np.int64)
        for o in range(world):
            f0, nf = int(fb[o]), int(fb[o + 1] - fb[o])
            if nf > 0:
                _arr(int(ptrs[o]) + 4 * rank * B * nf, B * nf, np.int32)[...] = xs[:, f0:f0 + nf].reshape(-1)
        return 0

    def dp_gather_push(self, recv_ids, off_local, shard, Vl, xs, out_bf16, ldx, col0, B, nf, E, world, packed, oob, s):
        if B == 0 or nf == 0:
            return 0
        ids = _arr(recv_ids, world * B * nf, np.int32).reshape(world, B, nf).astype(np.int64) + _arr(off_local, nf, np.int64)[None, None, :]
        ok = (ids >= 0) & (ids < Vl)
        tab = _mat(shard, Vl, E, E)
        rows = np.where(ok[..., None], tab[np.clip(ids, 0, Vl - 1)], F32(0)).reshape(world, B, nf * E)
        if oob and not ok.all():
            _arr(oob, 1, np.int32)[0] = 1
        ptrs = _arr(xs, world, np.int64)
        esz = 2 if out_bf16 else 4
        for p in range(world):
            base, ld = (int(ptrs[p]) + esz * B * col0, nf * E) if packed else (int(ptrs[p]) + esz * col0, ldx)
            if out_bf16:
                _mat(base, B, nf * E, ld, 1, np.uint16)[...] = f32_to_bf16(rows[p]).reshape(B, nf * E)
            else:
                _mat(base, B, nf * E, ld)[...] = rows[p]
        return 0

    def dp_push_grads(self, dX, ldg, B, F, E, grad_recv, out_bf16, fbound, rank, world, s):
        if B == 0:
            return 0
        fb = _arr(fbound, world + 1, np.int32)
        ptrs = _arr(grad_recv, world, np.int64)
        for o in range(world):
            f0, nf = int(fb[o]), int(fb[o + 1] - fb[o])
            if nf <= 0 or B == 0:
                continue
            g = _mat(dX + 4 * f0 * E, B, nf * E, ldg)
            if out_bf16:
                _mat(int(ptrs[o]) + 2 * rank * B * nf * E, B, nf * E, nf * E, 1, np.uint16)[...] = f32_to_bf16(g).reshape(B, nf * E)
            else:
                _mat(int(ptrs[o]) + 4 * rank * B * nf * E, B, nf * E, nf * E)[...] = g
        return 0

    # ---- a5/a6: first CGC level of PLE chained in one kernel (cdcmdr_ple_chain_fwd; ple.py:54,96-124, layer.py:184-190)
    def ple_chain_ok(self, K0, d0, d1, n_g):
        return 1 if (K0 >= 8 and K0 % 8 == 0 and -(-K0 // 64) <= 6 and d0 in (128, 256) and d1 in (64, 128) and 0 <= n_g <= 128) else 0

    def ple_chain_profile(self, counters):
        return 0

    def ple_chain_fwd(self, ref, s):
        p = _obj(ref)
        if p.drop_p > 0:
            raise NotImplementedError("the emulator does not reproduce the dropout hash; test with dropout=0")
        B, K0, nE, d0, d1, ng = int(p.B), int(p.K0), int(p.nE), int(p.d0), int(p.d1), int(p.n_g)
        if B == 0:
            return 0
        X = bf16_to_f32(_mat(p.X, B, K0, p.ldx, 1, np.uint16)).astype(np.float64)
        W0 = bf16_to_f32(_mat(p.W0, nE * d0 + ng, K0, K0, 1, np.uint16)).astype(np.float64)
        b0 = _arr(p.b0, nE * d0 + ng, np.float32)[:nE * d0 + ng]
        W1 = bf16_to_f32(_mat(p.W1, nE * d1, d0, d0, 1, np.uint16)).astype(np.float64)
        b1 = _arr(p.b1, nE * d1, np.float32)[:nE * d1]
        Z0 = (X @ W0.T).astype(np.float32) + b0[None, :]
        A0 = f32_to_bf16(np.maximum(Z0[:, :nE * d0], F32(0))).reshape(B, nE * d0)      # the second GEMM reads the bf16-rounded tile
        if p.A0:
            _mat(p.A0, B, nE * d0, p.lda0, 1, np.uint16)[...] = A0
        if ng:
            _mat(p.Lg, B, ng, p.ldg)[...] = Z0[:, nE * d0:]
        A0f = bf16_to_f32(A0).astype(np.float64)
        H = _mat(p.H, B, nE * d1, p.ldh, 1, np.uint16)
        for e in range(nE):
            z1 = (A0f[:, e * d0:(e + 1) * d0] @ W1[e * d1:(e + 1) * d1].T).astype(np.float32) + b1[None, e * d1:(e + 1) * d1]
            H[:, e * d1:(e + 1) * d1] = f32_to_bf16(np.maximum(z1, F32(0))).reshape(B, d1)
        return 0

    # ---- N4: AUC (midrank Mann-Whitney) + log loss per domain (cdcmdr_auc_logloss; run.py:677-711)
    def auc_logloss_scratch_bytes(self, n, n_domain):
        return 256

    def auc_logloss(self, pred, target, target_is_f32, domain, domain_is_i64, n, n_domain, out, scratch, s):
        o = _arr(out, n_domain * 4, np.float64)[:n_domain * 4].reshape(n_domain, 4)
        o[...] = np.nan
        if n == 0:
            return 0
        p = _arr(pred, n, np.float32)[:n]
        y = _arr(target, n, np.float32 if target_is_f32 else np.int16)[:n].astype(np.float32)
        d = _arr(domain, n, np.int64 if domain_is_i64 else np.int32)[:n].astype(np.int64) if domain else np.zeros(n, np.int64)
        eps = np.finfo(np.float32).eps
        pc = np.clip(p, eps, np.float32(1) - eps).astype(np.float64)
        for k in range(n_domain):
            m = d == k
            cnt = int(m.sum())
            pos = y[m] > 0
            npos = int(pos.sum())
            o[k, 2], o[k, 3] = npos, cnt
            if npos == 0 or npos == cnt:
                continue
            pk = p[m]
            order = np.argsort(pk, kind="stable")
            sp = pk[order]
            starts = np.flatnonzero(np.concatenate([[True], sp[1:] != sp[:-1]]))
            lens = np.diff(np.concatenate([starts, [cnt]]))
            mid = np.repeat(starts + 0.5 * (lens + 1), lens)                    # 1-based midranks in sorted order
            ranks = np.empty(cnt)
            ranks[order] = mid
            o[k, 0] = (ranks[pos].sum() - npos * (npos + 1) / 2.0) / (npos * (cnt - npos))
            o[k, 1] = -(np.log(pc[m][pos]).sum() + np.log(1.0 - pc[m][~pos]).sum()) / cnt
        return 0
