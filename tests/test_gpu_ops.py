"""Per-kernel parity on the GPU: every entry point of libcdcmdr.so is called through the C-ABI on device buffers and
compared with the host-memory restatement of the same entry point (oracle/host_abi.py) on identical seeded inputs.
Integer / index work must be bit-exact; fp32 work within 1e-5 of the tensor scale (summation order differs)."""
import ctypes as C

import numpy as np
import pytest
import torch

import cdcmdr_b200 as cm
from oracle.host_abi import HostABI

pytestmark = pytest.mark.gpu
L = cm._lib


def real():
    lib = L.Lib()
    return lib


class Env:
    """Runs the same call on (emulator, cpu) and (library, cuda) with identically seeded inputs."""

    def __init__(self, seed=0):
        self.seed = seed

    def run(self, fn):
        outs = []
        for lib, dev in ((HostABI(), "cpu"), (real(), "cuda")):
            self.rng = np.random.default_rng(self.seed)
            self.dev = dev
            self.keep = []
            res = fn(lib, self)
            if dev == "cuda":
                torch.cuda.synchronize()
            outs.append([r.detach().cpu().numpy().copy() for r in res])
        return outs

    def f32(self, *shape, scale=1.0):
        t = torch.from_numpy((self.rng.standard_normal(shape) * scale).astype(np.float32)).to(self.dev)
        self.keep.append(t)
        return t

    def zeros(self, *shape, dtype=torch.float32):
        t = torch.zeros(*shape, dtype=dtype, device=self.dev)
        self.keep.append(t)
        return t

    def ints(self, lo, hi, shape, dtype=np.int32):
        t = torch.from_numpy(self.rng.integers(lo, hi, size=shape).astype(dtype)).to(self.dev)
        self.keep.append(t)
        return t

    def put(self, arr):
        t = torch.from_numpy(np.ascontiguousarray(arr)).to(self.dev)
        self.keep.append(t)
        return t

    def scratch(self, nbytes):
        return self.zeros(max(int(nbytes), 256), dtype=torch.uint8)


def check(a, b, tol=1e-5, exact=False, what=""):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if exact:
        assert np.array_equal(a, b), f"{what}: not bit-exact ({(a != b).sum()} of {a.size} differ)"
        return
    scale = max(float(np.abs(a).max()), 1e-30)
    err = float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max())
    assert err <= tol * scale + 1e-7, f"{what}: max err {err:.3e} vs scale {scale:.3e}"


def both(fn, tol=1e-5, exact=False, seed=0):
    cpu, gpu = Env(seed).run(fn)
    assert len(cpu) == len(gpu)
    for i, (a, b) in enumerate(zip(cpu, gpu)):
        check(a, b, tol, exact if not isinstance(exact, (list, tuple)) else exact[i], what=f"output {i}")


# ------------------------------------------------------------------------------------------------ embedding
@pytest.mark.parametrize("B,F,E", [(1, 1, 4), (257, 16, 16), (1000, 23, 16), (64, 5, 3), (2048, 26, 32), (33, 7, 64)])
def test_embed_gather_bit_exact(B, F, E):
    fd = np.arange(3, 3 + F) * 7
    off = np.concatenate([[0], np.cumsum(fd)[:-1]]).astype(np.int64)
    V = int(fd.sum())

    def fn(lib, e):
        x = e.put(np.stack([e.rng.integers(0, d, size=B) for d in fd], axis=1).astype(np.int32))
        table = e.f32(V, E)
        out = e.zeros(B, F * E)
        ob = e.zeros(B, F * E + 8, dtype=torch.int16)
        oob = e.zeros(1, dtype=torch.int32)
        lib.embed_gather_fwd(x.data_ptr(), e.put(off).data_ptr(), table.data_ptr(), out.data_ptr(), ob.data_ptr(), F * E + 8,
                             B, F, E, V, oob.data_ptr(), 0)
        return [out, ob, oob]
    both(fn, exact=True)


def test_embed_gather_out_of_range_rows_read_zero():
    def fn(lib, e):
        x = e.put(np.array([[0, 5], [9, -1], [2, 1]], dtype=np.int32))
        off = e.put(np.array([0, 4], dtype=np.int64))
        table = e.f32(6, 8)
        out = e.zeros(3, 16)
        oob = e.zeros(1, dtype=torch.int32)
        lib.embed_gather_fwd(x.data_ptr(), off.data_ptr(), table.data_ptr(), out.data_ptr(), None, 0, 3, 2, 8, 6, oob.data_ptr(), 0)
        return [out, oob]
    both(fn, exact=True)


@pytest.mark.parametrize("B,F,E,V,skew", [(64, 4, 16, 50, False), (3000, 16, 16, 2000, True), (5000, 23, 16, 100000, True),
                                            (777, 6, 3, 40, True), (4096, 8, 64, 300, True)])
@pytest.mark.parametrize("mode", ["dense_grad", "adam_dense", "adam_dense_g16", "adam_lazy"])
def test_embed_backward(B, F, E, V, skew, mode):
    if mode == "adam_dense_g16" and E % 4:
        pytest.skip("bf16 row gradients need embed_dim % 4 == 0")
    per = V // F
    off = (np.arange(F) * per).astype(np.int64)

    def fn(lib, e):
        if skew:      # Zipf-like: a few rows collect thousands of entries -> long-segment kernel
            raw = np.minimum((e.rng.pareto(0.9, size=(B, F))).astype(np.int64), per - 1)
        else:
            raw = e.rng.integers(0, per, size=(B, F))
        x = e.put(raw.astype(np.int32))
        offs = e.put(off)
        ldg = F * E + 4
        go = e.f32(B, ldg)
        table, m, v = e.f32(V, E), e.f32(V, E, scale=0.01), e.put(np.abs(e.rng.standard_normal((V, E)).astype(np.float32)) * 0.01)
        st = e.zeros(48, dtype=torch.uint8)
        lib.step_state_init(st.data_ptr(), 0, 0)
        lib.step_tick(st.data_ptr(), 1e-3, 0.9, 0.99, 1e-8, 1e-8, 7, 0)
        lib.step_tick(st.data_ptr(), 1e-3, 0.9, 0.99, 1e-8, 1e-8, 7, 0)
        nb = lib.embed_plan_bytes(B * F, V, E)
        plan = e.scratch(nb)
        lib.embed_plan_build(x.data_ptr(), offs.data_ptr(), B, F, V, E, plan.data_ptr(), plan.numel(), 0)
        if mode == "dense_grad":
            g = e.zeros(V, E)
            lib.embed_bwd_dense(go.data_ptr(), ldg, plan.data_ptr(), E, B, F, E, V, g.data_ptr(), 0)
            return [g]
        if mode == "adam_dense":
            ss = e.zeros(1, dtype=torch.float64)
            lib.embed_bwd_adam_dense_exact(go.data_ptr(), ldg, plan.data_ptr(), E, B, F, E, V, table.data_ptr(), m.data_ptr(),
                                           v.data_ptr(), 1e-3, st.data_ptr(), ss.data_ptr(), 0)
            return [table, m, v, ss]
        if mode == "adam_dense_g16":                               # the replicas' exchange delivers bf16 row gradients
            ss = e.zeros(1, dtype=torch.float64)
            go16 = go.to(torch.bfloat16)
            e.keep.append(go16)
            lib.embed_bwd_adam_dense_exact_g16(go16.data_ptr(), ldg, plan.data_ptr(), E, B, F, E, V, table.data_ptr(), m.data_ptr(),
                                               v.data_ptr(), 1e-3, st.data_ptr(), ss.data_ptr(), 0)
            return [table, m, v, ss]
        lib.embed_bwd_adam_sparse_lazy(go.data_ptr(), ldg, plan.data_ptr(), E, B, F, E, V, table.data_ptr(), m.data_ptr(),
                                       v.data_ptr(), 1e-3, st.data_ptr(), 0)
        return [table, m, v]
    both(fn, tol=2e-5)


# ------------------------------------------------------------------------------------------------ fp32 GEMM
GEMM_CASES = [
    # M, N, K, G, layout, extras
    (37, 19, 23, 1, "nt", {}),
    (128, 64, 64, 3, "nt", dict(bias=True, act=1)),
    (300, 130, 70, 2, "nt", dict(bias=True, mask=True, accumulate=True)),
    (64, 48, 4096, 2, "tn", dict(split_k=8)),            # weight-gradient layout, K = batch
    (1, 32, 5000, 4, "tn", dict(split_k=5)),
    (513, 33, 1, 3, "nn", {}),
    (65, 1, 40, 4, "nt", dict(bias=True)),
]


@pytest.mark.parametrize("M,N,K,G,layout,extra", GEMM_CASES)
def test_gemm_f32(M, N, K, G, layout, extra):
    def fn(lib, e):
        if layout == "nt":     # A [G? ...] row-major [M, G*K]; Bt [G, N, K]
            A = e.f32(M, G * K); a_rs, a_cs, a_gs = G * K, 1, K
            Bt = e.f32(G, N, K); b_rs, b_cs, b_gs = K, 1, N * K
        elif layout == "tn":   # A(m,k) = dY[k, g*M + m] ; Bt(n,k) = X[k, g*N + n]
            A = e.f32(K, G * M); a_rs, a_cs, a_gs = 1, G * M, M
            Bt = e.f32(K, G * N); b_rs, b_cs, b_gs = 1, G * N, N
        else:                  # nn: Bt(n,k) = W[g, k, n]
            A = e.f32(M, G * K); a_rs, a_cs, a_gs = G * K, 1, K
            Bt = e.f32(G, K, N); b_rs, b_cs, b_gs = 1, N, N * K
        Cm = e.f32(M, G * N)
        bias = e.f32(G, N) if extra.get("bias") else None
        mask = e.f32(M, G * N) if extra.get("mask") else None
        split = extra.get("split_k", 1)
        wsp = e.zeros(split * G * M * N) if split > 1 else None
        d = L.GemmF32(A.data_ptr(), Bt.data_ptr(), Cm.data_ptr(), M, N, K, a_rs, a_cs, b_rs, b_cs, G * N, G, a_gs, b_gs, N,
                      bias.data_ptr() if bias is not None else None, N, extra.get("act", 0),
                      mask.data_ptr() if mask is not None else None, G * N, N, 1.25, 0.0, None, 0,
                      1 if extra.get("accumulate") else 0, split, wsp.data_ptr() if wsp is not None else None)
        lib.gemm_f32(C.byref(d), 0)
        return [Cm]
    both(fn, tol=2e-5)


# ------------------------------------------------------------------------------------------------ gate mix
@pytest.mark.parametrize("fast", [False, True])
@pytest.mark.parametrize("B,h,cfg", [(130, 128, "ple"), (77, 64, "ple_last"), (300, 64, "mmoe"), (64, 8, "mmoe"), (5, 6, "odd")])
def test_gate_mix_fwd_bwd(B, h, cfg, fast):
    """fast=True hands the launcher the pair count, which selects the row-per-warp backward kernel where the shape allows it
    (h in {64, 128}, <= 32 pairs); both kernels must agree with the emulator."""
    T, ns, nsh = 4, 2, 2
    nE = T * ns + nsh
    if cfg == "ple":
        ng, ms = T + 1, nE
        col = [t * 4 for t in range(T)] + [16]
        n = [4] * T + [nE]
        sel = sum(([*range(t * ns, (t + 1) * ns), *range(T * ns, nE)] + [0] * (ms - 4) for t in range(T)), []) + list(range(nE))
        ncol = 16 + nE + 1
    elif cfg == "ple_last":
        ng, ms = T, 4
        col = [t * 4 for t in range(T)]; n = [4] * T
        sel = sum(([*range(t * ns, (t + 1) * ns), *range(T * ns, nE)] for t in range(T)), [])
        ncol = 16
    elif cfg == "mmoe":
        nE, ng, ms = 8, 3, 8
        col = [0, 8, 16]; n = [8] * 3; sel = list(range(8)) * 3; ncol = 25
    else:
        nE, ng, ms = 3, 2, 3
        col = [1, 4]; n = [2, 3]; sel = [2, 0, 0, 0, 1, 2]; ncol = 7

    def fn(lib, e):
        desc = e.put(np.array(col + n + sel, dtype=np.int32))
        d = L.MixDesc(ng, nE, h, ms, desc.data_ptr(), desc.data_ptr() + 4 * ng, desc.data_ptr() + 8 * ng, sum(n) if fast else 0)
        H = torch.relu(e.f32(B, nE * h))
        logits = e.f32(B, ncol)
        out = e.zeros(B, ng * h)
        probs = e.zeros(B, ng * ms)
        lib.gate_mix_fwd(C.byref(d), H.data_ptr(), nE * h, logits.data_ptr(), ncol, out.data_ptr(), ng * h, probs.data_ptr(), B, 0, 0)
        dOut = e.f32(B, ng * h)
        dH = e.zeros(B, nE * h)
        dl = e.zeros(B, ncol)
        lib.gate_mix_bwd(C.byref(d), H.data_ptr(), nE * h, probs.data_ptr(), dOut.data_ptr(), ng * h, dH.data_ptr(), nE * h, 1.25,
                         dl.data_ptr(), ncol, B, 0, 0)
        dH2 = e.zeros(B, nE * h)
        dl2 = e.zeros(B, ncol)
        lib.gate_mix_bwd(C.byref(d), H.data_ptr(), nE * h, probs.data_ptr(), dOut.data_ptr(), ng * h, dH2.data_ptr(), nE * h, 0.0,
                         dl2.data_ptr(), ncol, B, 0, 0)
        return [out, probs, dH, dl, dH2]
    both(fn, tol=1e-5)


def _mix_cfg(cfg):
    T, ns, nsh = 4, 2, 2
    nE = T * ns + nsh
    if cfg == "ple":
        ng, ms = T + 1, nE
        col = [t * 4 for t in range(T)] + [16]
        n = [4] * T + [nE]
        sel = sum(([*range(t * ns, (t + 1) * ns), *range(T * ns, nE)] + [0] * (ms - 4) for t in range(T)), []) + list(range(nE))
        ncol = 16 + nE + 1
    elif cfg == "ple_last":
        ng, ms = T, 4
        col = [t * 4 for t in range(T)]; n = [4] * T
        sel = sum(([*range(t * ns, (t + 1) * ns), *range(T * ns, nE)] for t in range(T)), [])
        ncol = 16
    elif cfg == "mmoe":
        nE, ng, ms = 8, 3, 8
        col = [0, 8, 16]; n = [8] * 3; sel = list(range(8)) * 3; ncol = 25
    elif cfg == "wide":                                 # 16 experts, 2 gates of 16: every fragment slot in use
        nE, ng, ms = 16, 2, 16
        col = [0, 16]; n = [16, 16]; sel = list(range(16)) + list(range(15, -1, -1)); ncol = 32
    else:
        nE, ng, ms = 3, 2, 3
        col = [1, 4]; n = [2, 3]; sel = [2, 0, 0, 0, 1, 2]; ncol = 7
    return nE, ng, ms, col, n, sel, ncol


@pytest.mark.parametrize("B,h,cfg", [(130, 128, "ple"), (77, 64, "ple_last"), (300, 64, "mmoe"), (1000, 128, "mmoe"), (513, 128, "wide"),
                                     (9, 64, "odd"), (4100, 128, "ple"), (64, 32, "mmoe")])
def test_gate_mix_bf16(B, h, cfg):
    """bf16 rows: h in {64, 128} with <= 8 gates / 16 experts / 32 pairs run on the mma.sync kernels (gate_mix_mma.cu), the rest
    on the SIMT kernels; both against the emulator (fp32 math on the bf16-rounded inputs, rounded once on output)."""
    from oracle.host_abi import f32_to_bf16
    nE, ng, ms, col, n, sel, ncol = _mix_cfg(cfg)

    def bf(e, r, c, relu=False):
        v = e.rng.standard_normal((r, c)).astype(np.float32)
        if relu:
            v = np.maximum(v, 0)
        t = torch.from_numpy(f32_to_bf16(v).reshape(r, c).view(np.int16)).to(e.dev)
        e.keep.append(t)
        return t

    def fn(lib, e):
        desc = e.put(np.array(col + n + sel, dtype=np.int32))
        d = L.MixDesc(ng, nE, h, ms, desc.data_ptr(), desc.data_ptr() + 4 * ng, desc.data_ptr() + 8 * ng, sum(n))
        H = bf(e, B, nE * h, relu=True)
        logits = e.f32(B, ncol, scale=2.0)
        out = e.zeros(B, ng * h, dtype=torch.int16)
        probs = e.zeros(B, ng * ms)
        lib.gate_mix_fwd(C.byref(d), H.data_ptr(), nE * h, logits.data_ptr(), ncol, out.data_ptr(), ng * h, probs.data_ptr(), B, 1, 0)
        dOut = bf(e, B, ng * h)
        dH = e.zeros(B, nE * h, dtype=torch.int16)
        dl = e.zeros(B, ncol)
        lib.gate_mix_bwd(C.byref(d), H.data_ptr(), nE * h, probs.data_ptr(), dOut.data_ptr(), ng * h, dH.data_ptr(), nE * h, 1.25,
                         dl.data_ptr(), ncol, B, 1, 0)
        dH2 = e.zeros(B, nE * h, dtype=torch.int16)
        dl2 = e.zeros(B, ncol)
        lib.gate_mix_bwd(C.byref(d), H.data_ptr(), nE * h, probs.data_ptr(), dOut.data_ptr(), ng * h, dH2.data_ptr(), nE * h, 0.0,
                         dl2.data_ptr(), ncol, B, 1, 0)
        return [out, probs, dH, dl, dH2, dl2]

    cpu, gpu = Env(3).run(fn)
    for i, (a, b) in enumerate(zip(cpu, gpu)):
        if a.dtype == np.int16:                          # bf16 payload: within one bf16 ulp of the tensor scale, almost all equal
            af = (a.view(np.uint16).astype(np.uint32) << 16).view(np.float32)
            bfv = (b.view(np.uint16).astype(np.uint32) << 16).view(np.float32)
            err = np.abs(af - bfv)
            assert float((err > 2 ** -7 * np.maximum(np.abs(af), 1e-3)).mean()) == 0.0, f"output {i}: more than one bf16 ulp off"
            assert float((a != b).mean()) < 0.01, f"output {i}: {(a != b).mean():.4f} of the bf16 entries differ"
        else:
            check(a, b, tol=2e-5, what=f"output {i}")


# ------------------------------------------------------------------------------------------------ batch norm
@pytest.mark.parametrize("B,Cn,train,relu,g2", [(257, 70, 1, 1, 0), (4096, 256, 1, 1, 0), (100, 33, 0, 1, 0), (64, 16, 1, 0, 1),
                                                  (2, 5, 1, 1, 0)])
def test_batch_norm_fwd_bwd(B, Cn, train, relu, g2):
    def fn(lib, e):
        ldz = Cn + 3
        Z = e.f32(B, ldz, scale=2.0)
        gam, bet = e.f32(Cn), e.f32(Cn)
        gam2, bet2 = (e.f32(Cn), e.f32(Cn)) if g2 else (None, None)
        rm, rv = e.f32(Cn), e.put(np.abs(e.rng.standard_normal(Cn)).astype(np.float32) + 0.5)
        sm, si = e.zeros(Cn), e.zeros(Cn)
        A = e.zeros(B, Cn)
        sc = e.scratch(lib.bn_scratch_bytes(Cn))
        d = L.BnDesc(gam.data_ptr(), bet.data_ptr(), gam2.data_ptr() if g2 else None, bet2.data_ptr() if g2 else None,
                     rm.data_ptr(), rv.data_ptr(), sm.data_ptr(), si.data_ptr(), train, relu, 0.0, None, 0)
        lib.bn_fwd(C.byref(d), Z.data_ptr(), ldz, A.data_ptr(), Cn, 0, B, Cn, sc.data_ptr(), 0)
        dA = e.f32(B, Cn)
        dZ = e.zeros(B, Cn)
        dg, db = e.f32(Cn), e.f32(Cn)
        lib.bn_bwd(C.byref(d), Z.data_ptr(), ldz, A.data_ptr(), Cn, 0, dA.data_ptr(), Cn, 0, dZ.data_ptr(), Cn, 0, dg.data_ptr(),
                   db.data_ptr(), 1, B, Cn, sc.data_ptr(), 0)
        return [A, sm, si, rm, rv, dZ, dg, db]
    both(fn, tol=3e-5)


# ------------------------------------------------------------------------------------------------ loss
@pytest.mark.parametrize("B,T,mode", [(1000, 4, 0), (1000, 4, 1), (1000, 4, 2), (513, 1, 1), (300, 3, 3)])
@pytest.mark.parametrize("tf32", [0, 1])
def test_sigmoid_select_bce(B, T, mode, tf32):
    if mode == 3 and tf32:
        pytest.skip("forward-only mode takes no target")

    def fn(lib, e):
        logits = e.f32(B, T, scale=3.0)
        lin = e.f32(B, 5)
        sel = e.ints(0, T, (B,), np.int64)
        tg = e.put((e.rng.random(B) < 0.3).astype(np.float32 if tf32 else np.int16))
        pred, psel = e.zeros(B, T), e.zeros(B)
        ls = e.zeros(1, dtype=torch.float64)
        dl = e.zeros(B, T)
        dlin = e.zeros(B, 3)
        sc = e.scratch(lib.reduce_scratch_bytes())
        has_t = mode != 3
        lib.sigmoid_select_bce(logits.data_ptr(), lin.data_ptr() + 8, 5, B, T, mode, sel.data_ptr(), 1 % T,
                               tg.data_ptr() if has_t else None, tf32, pred.data_ptr(), psel.data_ptr(),
                               ls.data_ptr() if has_t else None, dl.data_ptr() if has_t else None,
                               dlin.data_ptr() + 4 if has_t else None, 3, 1.0 / B, sc.data_ptr(), 0)
        dp = e.f32(B, T)
        dl2, dlin2 = e.zeros(B, T), e.zeros(B)
        lib.sigmoid_bwd(pred.data_ptr(), dp.data_ptr(), dl2.data_ptr(), dlin2.data_ptr(), 1, B, T, 0)
        return [pred, psel, ls, dl, dlin, dl2, dlin2]
    both(fn, tol=2e-5)


def test_bce_saturated_probabilities_are_clamped():
    def fn(lib, e):
        logits = e.put(np.array([[200.0], [-200.0], [200.0], [-200.0], [0.0]], dtype=np.float32))
        tg = e.put(np.array([0, 1, 1, 0, 1], dtype=np.int16))
        pred, ls, dl = e.zeros(5, 1), e.zeros(1, dtype=torch.float64), e.zeros(5, 1)
        sc = e.scratch(lib.reduce_scratch_bytes())
        lib.sigmoid_select_bce(logits.data_ptr(), None, 0, 5, 1, 1, None, 0, tg.data_ptr(), 0, pred.data_ptr(), None, ls.data_ptr(),
                               dl.data_ptr(), None, 0, 0.2, sc.data_ptr(), 0)
        return [pred, ls, dl]
    both(fn, tol=1e-5)


def test_out_of_range_tower_selection_is_loud():
    """mode 0 with a selection outside [0, T) (domain_to_group gives -1 for an unknown domain id): the reference's y_cat.gather
    raises; here the row's prediction and the loss are NaN and the row contributes no gradient."""
    def fn(lib, e):
        logits = e.f32(6, 3)
        sel = e.put(np.array([0, 2, -1, 1, 3, 2], dtype=np.int64))
        tg = e.put(np.array([0, 1, 1, 0, 1, 0], dtype=np.int16))
        pred, ps, ls, dl = e.zeros(6, 3), e.zeros(6), e.zeros(1, dtype=torch.float64), e.zeros(6, 3)
        sc = e.scratch(lib.reduce_scratch_bytes())
        lib.sigmoid_select_bce(logits.data_ptr(), None, 0, 6, 3, 0, sel.data_ptr(), 0, tg.data_ptr(), 0, pred.data_ptr(), ps.data_ptr(),
                               ls.data_ptr(), dl.data_ptr(), None, 0, 1 / 6, sc.data_ptr(), 0)
        return [pred, ps, ls, dl]
    cpu, gpu = Env(3).run(fn)
    for out in (cpu, gpu):
        pred, ps, ls, dl = out
        assert np.isnan(ps[[2, 4]]).all() and not np.isnan(ps[[0, 1, 3, 5]]).any()
        assert np.isnan(ls[0])
        assert not dl[[2, 4]].any() and not np.isnan(dl).any() and np.abs(dl[[0, 1, 3, 5]]).sum() > 0
    check(cpu[0], gpu[0], tol=1e-6)
    check(cpu[3], gpu[3], tol=1e-5)


# ------------------------------------------------------------------------------------------------ regulariser / Adam / reductions
def test_reg_adam_colsum_and_elementwise():
    n = 100003

    def fn(lib, e):
        w, g, m = e.f32(n), e.f32(n, scale=0.1), e.f32(n, scale=0.01)
        v = e.put(np.abs(e.rng.standard_normal(n)).astype(np.float32) * 0.01)
        coef = e.put((e.rng.random(n) < 0.5).astype(np.float32) * 1e-3)
        present = e.put((e.rng.random(n) < 0.9).astype(np.uint8))
        st = e.zeros(48, dtype=torch.uint8)
        lib.step_state_init(st.data_ptr(), 4, 0)
        lib.step_tick(st.data_ptr(), 1e-3, 0.9, 0.99, 1e-8, 1e-8, 11, 0)
        sc = e.scratch(lib.reduce_scratch_bytes())
        o1, o2 = e.zeros(1, dtype=torch.float64), e.zeros(1, dtype=torch.float64)
        lib.reg_l2_sum(w.data_ptr(), coef.data_ptr(), 0.0, n, o1.data_ptr(), sc.data_ptr(), 0)
        lib.reg_l2_sum(w.data_ptr(), None, 0.5, n, o2.data_ptr(), sc.data_ptr(), 0)
        rg = e.f32(n)
        lib.reg_l2_grad(w.data_ptr(), coef.data_ptr(), 0.0, 0.7, rg.data_ptr(), 1, n, 0)
        rg2 = e.zeros(n)
        lib.reg_l2_grad(w.data_ptr(), None, 0.25, 1.0, rg2.data_ptr(), 0, n, 0)
        lib.adam_dense(w.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), coef.data_ptr(), present.data_ptr(), n, st.data_ptr(), 0)
        B, Cn = 3001, 77
        X = e.f32(B, Cn + 5)
        cs = e.f32(Cn)
        csc = e.scratch(lib.colsum_scratch_bytes(Cn))
        lib.colsum(X.data_ptr() + 8, Cn + 5, 0, B, Cn, cs.data_ptr(), 1, csc.data_ptr(), 0)
        a, b, o = e.f32(n), e.f32(n), e.f32(n)
        outs = []
        for op in range(4):
            oo = o.clone(); e.keep.append(oo)
            lib.ewise_f32(a.data_ptr(), b.data_ptr(), oo.data_ptr(), n, op, 0)
            outs.append(oo)
        rm = e.zeros(B, Cn)
        lib.relu_mask_f32(X.data_ptr(), Cn + 5, X.data_ptr() + 4, Cn + 5, rm.data_ptr(), Cn, B, Cn, 1.25, 0)
        ad = e.f32(B, Cn)
        lib.add2d_f32(X.data_ptr(), Cn + 5, ad.data_ptr(), Cn, B, Cn, 1, 0)
        bf = e.zeros(B, Cn + 1, dtype=torch.int16)
        lib.cast_f32_bf16(X.data_ptr(), Cn + 5, bf.data_ptr(), Cn + 1, B, Cn, 0)
        back = e.f32(B, Cn)
        lib.cast_bf16_f32(bf.data_ptr(), Cn + 1, back.data_ptr(), Cn, B, Cn, 1, 0)
        return [o1, o2, rg, rg2, w, m, v, cs, *outs, rm, ad, bf, back]
    cpu, gpu = Env(3).run(fn)
    for i, (a, b) in enumerate(zip(cpu, gpu)):
        check(a, b, tol=1e-5, exact=(i == len(cpu) - 2), what=f"output {i}")


def test_step_tick_matches_torch_adam_bias_corrections():
    def fn(lib, e):
        st = e.zeros(48, dtype=torch.uint8)
        lib.step_state_init(st.data_ptr(), 0, 0)
        outs = []
        for _ in range(3):
            lib.step_tick(st.data_ptr(), 1e-3, 0.9, 0.99, 1e-8, 1e-8, 5, 0)
            outs.append(st.clone())
        return outs
    cpu, gpu = Env().run(fn)
    for t, (a, b) in enumerate(zip(cpu, gpu)):
        assert np.array_equal(a[:8], b[:8])                      # step counter
        fa, fb = a[16:40].view(np.float32), b[16:40].view(np.float32)
        np.testing.assert_allclose(fa, fb, rtol=1e-6)
        assert abs(fb[0] - 1e-3 / (1 - 0.9 ** (t + 1))) < 1e-8 and abs(fb[5] - np.sqrt(1 - 0.99 ** (t + 1))) < 1e-6
    assert len({bytes(g[8:16]) for g in gpu}) == 3               # the dropout seed changes every step


# ------------------------------------------------------------------------------------------------ cross / softmax / tanh
@pytest.mark.parametrize("B,D", [(100, 24), (257, 832), (1, 7)])
def test_cross_ops(B, D):
    def fn(lib, e):
        x0, x, b = e.f32(B, D), e.f32(B, D), e.f32(D)
        xw1, xwD = e.f32(B), e.f32(B, D)
        o1, o2 = e.zeros(B, D), e.zeros(B, D)
        lib.cross_fuse_fwd(x0.data_ptr(), x.data_ptr(), xw1.data_ptr(), 1, b.data_ptr(), o1.data_ptr(), B, D, 0)
        lib.cross_fuse_fwd(x0.data_ptr(), x.data_ptr(), xwD.data_ptr(), D, b.data_ptr(), o2.data_ptr(), B, D, 0)
        do = e.f32(B, D)
        acc1, acc2 = e.f32(B, D), e.f32(B, D)
        dxw1, dxwD = e.zeros(B), e.zeros(B, D)
        lib.cross_fuse_bwd(x0.data_ptr(), xw1.data_ptr(), 1, do.data_ptr(), acc1.data_ptr(), dxw1.data_ptr(), B, D, 0)
        lib.cross_fuse_bwd(x0.data_ptr(), xwD.data_ptr(), D, do.data_ptr(), acc2.data_ptr(), dxwD.data_ptr(), B, D, 0)
        ne = 3
        u, g = e.f32(ne, B, D), torch.softmax(e.f32(B, ne), dim=1)
        e.keep.append(g)
        om = e.zeros(B, D)
        lib.crossmix_combine_fwd(x0.data_ptr(), x.data_ptr(), u.data_ptr(), g.data_ptr(), b.data_ptr(), om.data_ptr(), B, D, ne, 0)
        du, dg, acc3 = e.zeros(ne, B, D), e.zeros(B, ne), e.f32(B, D)
        lib.crossmix_combine_bwd(x0.data_ptr(), u.data_ptr(), g.data_ptr(), b.data_ptr(), do.data_ptr(), du.data_ptr(), dg.data_ptr(),
                                 acc3.data_ptr(), B, D, ne, 0)
        t = e.f32(B, D)
        lib.tanh_fwd(t.data_ptr(), B * D, 0)
        dt = e.f32(B, D)
        lib.tanh_bwd(t.data_ptr(), dt.data_ptr(), B * D, 0)
        z = e.f32(B, 9)
        p, dz = e.zeros(B, 4), e.zeros(B, 4)
        lib.softmax_rows_fwd(z.data_ptr() + 4, 9, p.data_ptr(), 4, B, 4, 0)
        dp = e.f32(B, 4)
        lib.softmax_rows_bwd(p.data_ptr(), 4, dp.data_ptr(), 4, dz.data_ptr(), 4, B, 4, 0)
        return [o1, o2, acc1, acc2, dxw1, dxwD, om, du, dg, acc3, t, dt, p, dz]
    both(fn, tol=2e-5)


# ------------------------------------------------------------------------------------------------ routing
@pytest.mark.parametrize("B,ng", [(12, 4), (1, 1), (1023, 3), (1024, 7), (1025, 256), (70000, 30), (5000, 5)])
def test_route_partition_bit_exact(B, ng):
    def fn(lib, e):
        if B == 12:
            g = np.array([0, 1, 0, 3, 1, 0, 2, 1, 3, 3, 0, 1], dtype=np.int64)       # SURVEY §3.5 probe
        elif B == 5000:
            g = e.rng.integers(-1, ng + 1, size=B).astype(np.int64)                  # out-of-range ids are dropped
            g[g == 2] = 4                                                            # an empty group
        else:
            g = e.rng.integers(0, ng, size=B).astype(np.int64)
        gt = e.put(g)
        perm = e.zeros(B, dtype=torch.int32) - 1
        e.keep.append(perm)
        cnt, gs = e.zeros(ng, dtype=torch.int32), e.zeros(ng + 1, dtype=torch.int32)
        sc = e.scratch(lib.route_scratch_bytes(B, ng))
        lib.route_partition(gt.data_ptr(), B, ng, perm.data_ptr(), cnt.data_ptr(), gs.data_ptr(), sc.data_ptr(), 0)
        n_routed = int(((g >= 0) & (g < ng)).sum())
        src = e.f32(B, 20)
        dst = e.zeros(max(n_routed, 1), 24)
        lib.permute_rows(src.data_ptr(), 20, perm.data_ptr(), n_routed, 20, 4, dst.data_ptr(), 24, 0, 0)
        back = e.zeros(B, 20)
        lib.permute_rows(dst.data_ptr(), 24, perm.data_ptr(), n_routed, 20, 4, back.data_ptr(), 20, 1, 0)
        return [perm, cnt, gs, dst, back]
    cpu, gpu = Env(5).run(fn)
    for a, b in zip(cpu, gpu):
        assert np.array_equal(a, b)
    if B == 12:
        assert gpu[0].tolist() == [0, 2, 5, 10, 1, 4, 7, 11, 6, 3, 8, 9]


def test_domain_to_group_bit_exact():
    def fn(lib, e):
        B, F, nd = 3333, 23, 30
        x = e.ints(0, nd, (B, F))
        d2g = e.ints(0, 4, (nd,), np.int64)
        out = e.zeros(B, dtype=torch.int64)
        lib.domain_to_group(x.data_ptr(), B, F, 10, d2g.data_ptr(), nd, out.data_ptr(), 0)
        return [out]
    both(fn, exact=True)


@pytest.mark.parametrize("batches,rows,cols,elt", [(4, 4, 128, 4), (3, 1, 5, 4), (5, 7, 33, 2), (1, 16, 64, 4)])
def test_copy2d_batched_bit_exact(batches, rows, cols, elt):
    """strided batch of 2-D copies (gate blocks -> block-diagonal operand and back)"""
    def fn(lib, e):
        lds, ldd = cols + 3, 2 * cols + 5
        src_bs, dst_bs = rows * lds + 7, rows * ldd + cols
        dt = torch.int32 if elt == 4 else torch.int16
        src = e.put(e.rng.integers(-1000, 1000, size=batches * src_bs + 16).astype(np.int32 if elt == 4 else np.int16))
        dst = e.zeros(batches * dst_bs + rows * ldd + 16, dtype=dt)
        lib.copy2d_batched(src.data_ptr(), src_bs, lds, dst.data_ptr(), dst_bs, ldd, batches, rows, cols, elt, 0)
        back = e.zeros(batches * src_bs + 16, dtype=dt)
        lib.copy2d_batched(dst.data_ptr(), dst_bs, ldd, back.data_ptr(), src_bs, lds, batches, rows, cols, elt, 0)
        return [dst, back]
    both(fn, exact=True)


@pytest.mark.parametrize("elt", [2, 4])
def test_copy2d_batched_aligned_blocks(elt):
    """the exchange blocks of the sharded table: every stride a multiple of 16 bytes (128-bit path), 3 owners x [B, n*E]"""
    B, nE, F_E = 257, 48, 368
    def fn(lib, e):
        dt = torch.int32 if elt == 4 else torch.int16
        npdt = np.int32 if elt == 4 else np.int16
        X = e.put(e.rng.integers(-1000, 1000, size=B * F_E).astype(npdt))
        packed = e.zeros(3 * B * nE, dtype=dt)
        lib.copy2d_batched(X.data_ptr() + elt * 16, nE, F_E, packed.data_ptr(), B * nE, nE, 3, B, nE, elt, 0)
        back = e.zeros(B * F_E, dtype=dt)
        lib.copy2d_batched(packed.data_ptr(), B * nE, nE, back.data_ptr() + elt * 16, nE, F_E, 3, B, nE, elt, 0)
        return [packed, back]
    both(fn, exact=True)
