"""tcgen05/TMA bf16 GEMM (cdcmdr_gemm_bf16_tc) against the host restatement on identical bf16 inputs.  The fp32
accumulation order differs (tensor core vs float64 dot), so fp32 outputs agree to ~1e-5 of the tensor scale and bf16
outputs to one bf16 ulp (2^-8 relative) of the tensor scale."""
import ctypes as C

import numpy as np
import pytest
import torch

import cdcmdr_b200 as cm
from oracle.host_abi import HostABI, f32_to_bf16
from tests.test_gpu_ops import Env, check

pytestmark = pytest.mark.gpu
L = cm._lib

# name: M, N, K, G, a_mn, b_mn, n_main(-1 = all main), extras
CASES = {
    "k_major_exact_tile": (128, 256, 64, 1, 0, 0, -1, {}),
    "k_major_k_tail": (300, 240, 368, 1, 0, 0, -1, dict(bias=True, act=1)),
    "k_major_multi_tile": (1000, 2587, 368, 1, 0, 0, 2560, dict(bias=True, act=1)),
    "k_major_aux_only": (513, 27, 128, 1, 0, 0, 0, dict(bias=True)),
    "grouped_k_offsets": (777, 128, 256, 5, 0, 0, -1, dict(bias=True, act=1, a_gk=256, b_gn=128, main_gn=128)),
    "mask_accumulate": (260, 368, 2587, 1, 0, 0, -1, dict(mask=True, accumulate=True)),
    "b_mn_major_dgrad": (640, 368, 2587, 1, 0, 1, -1, {}),
    "grouped_b_mn_major": (400, 256, 128, 3, 0, 1, -1, dict(mask=True, a_gk=128, b_gk=128, main_gn=256)),
    "wgrad_both_mn_major": (2587, 368, 4096, 1, 1, 1, 0, {}),
    "wgrad_split_k": (200, 130, 8192, 1, 1, 1, 0, dict(split_k=6)),
    # long split-K slices (>= 48 k-blocks each): the lean-epilogue launch (192 threads) - plain, ragged, and grouped
    "wgrad_lean": (300, 376, 16384, 1, 1, 1, 0, dict(split_k=4)),
    "wgrad_lean_ragged": (2587, 376, 8192 + 64 * 5 + 24, 1, 1, 1, 0, dict(split_k=2)),
    "wgrad_lean_grouped": (128, 256, 8192, 3, 1, 1, 0, dict(a_gm=128, b_gn=256, aux_gn=256 * 128, split_k=2, group_rows=True)),
    "wgrad_grouped": (128, 256, 2048, 4, 1, 1, 0, dict(a_gm=128, b_gn=256, aux_gn=256 * 128, split_k=3, group_rows=True)),
    "tiny": (5, 16, 8, 1, 0, 0, -1, {}),
    "pair_odd_row_tiles": (128 * 5 + 17, 256, 200, 1, 0, 0, -1, dict(bias=True, act=1)),
    "pair_aux_f32": (1024, 320, 96, 1, 0, 0, 0, dict(bias=True)),
    "pair_many_tiles": (128 * 40, 512, 64, 1, 0, 0, -1, {}),
    # A-resident column sweep: short K, several column tiles - ungrouped with an odd tile count per CTA range, and grouped
    "sweep_ragged_ranges": (128 * 37 + 5, 3 * 256 + 40, 368, 1, 0, 0, -1, dict(bias=True, act=1)),
    "sweep_grouped": (500, 304, 128, 3, 0, 0, -1, dict(bias=True, act=1, a_gk=128, b_gn=304, main_gn=304)),
}


@pytest.fixture(params=[4, 1, 3, 0, 9], ids=["cta_pairs", "single_cta", "single_cta_streamed_a", "auto", "single_cta_full_epilogue"])
def tc_mode(request):
    """cdcmdr_gemm_bf16_tc tile mode: CTA pairs (cta_group::2) where the shape allows / single-CTA tiles only."""
    lib = cm._lib.load()
    old = lib.gemm_bf16_tc_mode(request.param)
    yield request.param
    lib.gemm_bf16_tc_mode(old)


@pytest.mark.parametrize("name", sorted(CASES))
def test_gemm_bf16_tc(name, tc_mode):
    M, N, K, G, a_mn, b_mn, n_main, ex = CASES[name]
    n_main = N if n_main < 0 else n_main

    def bf(e, r, c, pad=0, scale=1.0):
        ld = (c + 7) // 8 * 8 + (8 if pad else 0)
        v = (e.rng.standard_normal((r, ld)) * scale).astype(np.float32)
        t = torch.from_numpy(f32_to_bf16(v).reshape(r, ld).view(np.int16)).to(e.dev)
        e.keep.append(t)
        return t, ld

    def fn(lib, e):
        a_gm, a_gk, b_gn, b_gk = ex.get("a_gm", 0), ex.get("a_gk", 0), ex.get("b_gn", 0), ex.get("b_gk", 0)
        if a_mn:
            A, lda = bf(e, K, max(M, a_gm * (G - 1) + M), pad=8)
            a_rows, a_cols = K, max(M, a_gm * (G - 1) + M)
        else:
            A, lda = bf(e, M, max(K, a_gk * (G - 1) + K), pad=8)
            a_rows, a_cols = M, max(K, a_gk * (G - 1) + K)
        if b_mn:
            Bt, ldb = bf(e, max(K, b_gk * (G - 1) + K), max(N, b_gn * (G - 1) + N), pad=0 if N % 8 == 0 else 8 - N % 8, scale=0.1)
            b_rows, b_cols = max(K, b_gk * (G - 1) + K), max(N, b_gn * (G - 1) + N)
        else:
            Bt, ldb = bf(e, max(N, b_gn * (G - 1) + N), K, pad=0 if K % 8 == 0 else 8 - K % 8, scale=0.1)
            b_rows, b_cols = max(N, b_gn * (G - 1) + N), K
        main_gn = ex.get("main_gn", 0)
        ld_main = max(n_main, main_gn * (G - 1) + n_main) + 8
        out_main, _ = bf(e, M, ld_main - 8, pad=8)
        n_aux = N - n_main
        aux_gn = ex.get("aux_gn", 0)
        want = ex.get("split_k", 1)
        split = lib.gemm_bf16_tc_splits(K, want) if want > 1 else 1
        if ex.get("group_rows"):                      # each group's [M, N] block stacked: aux_gn = M*N, ld = N
            ld_aux = n_aux
            aux_elems = G * M * n_aux
        else:
            ld_aux = max(n_aux, 1) + 4
            aux_elems = M * ld_aux
        out_aux = e.f32(max(split, 1) * aux_elems)
        bias = e.f32(G, N) if ex.get("bias") else None
        mask, ld_mask = (bf(e, M, ld_main - 8, pad=8)) if ex.get("mask") else (None, 0)
        d = L.GemmBf16(A.data_ptr(), lda, a_rows, a_cols, Bt.data_ptr(), ldb, b_rows, b_cols, M, N, K, G, a_gm, a_gk, b_gn, b_gk,
                       a_mn, b_mn, bias.data_ptr() if bias is not None else None, N, n_main,
                       out_main.data_ptr() if n_main else None, ld_main, main_gn,
                       out_aux.data_ptr() if n_aux else None, ld_aux, aux_gn, ex.get("act", 0),
                       mask.data_ptr() if mask is not None else None, ld_mask, main_gn, 1.25, 0.0, None, 0,
                       1 if ex.get("accumulate") else 0, split, aux_elems, 0)
        lib.gemm_bf16_tc(C.byref(d), 0)
        outs = []
        if n_main:
            outs.append(out_main)
        if n_aux:
            if split > 1:
                red = e.zeros(aux_elems)
                lib.splitk_reduce(out_aux.data_ptr(), aux_elems, split, red.data_ptr(), 1, aux_elems, aux_elems, aux_elems, 0, 0)
                outs.append(red)
            else:
                outs.append(out_aux)
        return outs

    cpu, gpu = Env(11).run(fn)
    i = 0
    if n_main:
        a = (cpu[i].view(np.uint16).astype(np.uint32) << 16).view(np.float32)
        b = (gpu[i].view(np.uint16).astype(np.uint32) << 16).view(np.float32)
        scale = max(float(np.abs(a).max()), 1e-30)
        err = float(np.abs(a - b).max())
        assert err <= 2 ** -7 * scale, f"{name} bf16 main: err {err:.3e} scale {scale:.3e}"
        assert float((np.abs(a - b) > 2 ** -9 * scale).mean()) < 0.02, f"{name}: too many bf16 entries differ"
        i += 1
    if n_main < N:
        check(cpu[i], gpu[i], tol=3e-5, what=f"{name} fp32 aux")


def test_transpose_bf16():
    def fn(lib, e):
        src = e.ints(-30000, 30000, (300, 77), np.int16)
        dst = e.zeros(77, 304, dtype=torch.int16)
        lib.transpose_bf16(src.data_ptr(), 77, dst.data_ptr(), 304, 300, 77, 0)
        return [dst]
    cpu, gpu = Env().run(fn)
    assert np.array_equal(cpu[0], gpu[0])


@pytest.mark.parametrize("ld_pad", [0, 4])
def test_gemm_bf16_tc_dropout_statistics(ld_pad):
    """nn.Dropout(p) in the epilogue (layer.py:189): with A = 0 and bias = 1 the output is exactly keep/(1-p): the keep rate must be
    1-p within sampling error, kept entries are scaled by 1/(1-p), the mask is a function of (seed, salt) only (bit-repeatable),
    changes with the salt, and has no row/column structure.  ld_pad=4 forces the non-TMA store path (pitch not 16-byte aligned)."""
    lib = cm._lib.load()
    dev = torch.device("cuda")
    M, N, K, p = 4096, 512, 64, 0.2
    A = torch.zeros(M, K, dtype=torch.bfloat16, device=dev)
    Bt = torch.zeros(N, K, dtype=torch.bfloat16, device=dev)
    bias = torch.ones(N, device=dev)
    st = torch.zeros(48, dtype=torch.uint8, device=dev)
    lib.step_state_init(st.data_ptr(), 3, 0)
    lib.step_tick(st.data_ptr(), 1e-3, 0.9, 0.99, 1e-8, 0.0, 2000, 0)
    ld = N + ld_pad

    def run(salt):
        out = torch.full((M, ld), -7.0, dtype=torch.bfloat16, device=dev)
        d = L.GemmBf16(A.data_ptr(), K, M, K, Bt.data_ptr(), K, N, K, M, N, K, 1, 0, 0, 0, 0, 0, 0, bias.data_ptr(), 0, N,
                       out.data_ptr(), ld, 0, None, 0, 0, 1, None, 0, 0, 1.0, p, st.data_ptr() + 8, salt, 0, 1, 0, 0)
        lib.gemm_bf16_tc(C.byref(d), 0)
        torch.cuda.synchronize()
        return out.float().cpu().numpy()
    a, b, c = run(5), run(5), run(6)
    assert np.array_equal(a, b)
    if ld_pad:
        assert (a[:, N:] == -7.0).all()                       # padding columns untouched
    a, c = a[:, :N], c[:, :N]
    assert set(np.unique(a).tolist()) == {0.0, 1.25}
    keep = (a > 0)
    assert abs(keep.mean() - (1 - p)) < 4e-3, keep.mean()
    assert abs((c > 0).mean() - (1 - p)) < 4e-3
    assert 0.25 < (keep != (c > 0)).mean() < 0.40              # independent masks differ on 2p(1-p) = 32% of the entries
    assert np.abs(keep.mean(0) - (1 - p)).max() < 0.04 and np.abs(keep.mean(1) - (1 - p)).max() < 0.09
    # neighbouring columns / rows are uncorrelated
    k = keep.astype(np.float64) - keep.mean()
    assert abs((k[:, 1:] * k[:, :-1]).mean()) < 2e-3 and abs((k[1:] * k[:-1]).mean()) < 2e-3


@pytest.mark.parametrize("M,D", [(1000, 832), (300, 96), (129, 256), (5, 64), (4096, 32)])
def test_gemm_bf16_tc_cross_epilogue(M, D, tc_mode):
    """CrossNetV2 layer fused into the GEMM epilogue (layer.py:339-343): y = x0 * (x W^T) + b + x with bf16 x0 / x boxes fetched by
    TMA, bf16 y (the next layer's operand) and the raw fp32 product (kept for the backward) stored by TMA - against the host
    restatement on identical bf16 operands."""
    def fn(lib, e):
        b16 = lambda a: e.put(f32_to_bf16(a.astype(np.float32)).reshape(a.shape).view(np.int16))   # noqa: E731
        xb = b16(e.rng.standard_normal((M, D)))
        x0b = b16(e.rng.standard_normal((M, D)))
        W = b16(e.rng.standard_normal((D, D)) * 0.05)
        bias = e.f32(D)
        xw = e.zeros(M, D)
        yb = e.zeros(M, D, dtype=torch.int16)
        d = L.GemmBf16(A=xb.data_ptr(), lda=D, a_rows=M, a_cols=D, Bt=W.data_ptr(), ldb=D, b_rows=D, b_cols=D, M=M, N=D, K=D, G=1,
                       bias=bias.data_ptr(), n_main=D, out_main=yb.data_ptr(), ld_main=D, out_aux=xw.data_ptr(), ld_aux=D,
                       mask_scale=1.0, split_k=1, cross_x0=x0b.data_ptr(), cross_x=xb.data_ptr(), ld_cross=D)
        lib.gemm_bf16_tc(C.byref(d), 0)
        return [xw, yb]
    cpu, gpu = Env(21).run(fn)
    check(cpu[0], gpu[0], tol=3e-5, what="cross acc")
    a = (cpu[1].view(np.uint16).astype(np.uint32) << 16).view(np.float32)
    b = (gpu[1].view(np.uint16).astype(np.uint32) << 16).view(np.float32)
    scale = float(np.abs(a).max())
    assert float(np.abs(a - b).max()) <= 2 ** -7 * scale


CHAIN_CASES = {
    # name: B, K0, nE, d0, d1, n_g, store_a0
    "c4_shape": (1000, 368, 10, 256, 128, 27, True),
    "c4_inference": (1000, 368, 10, 256, 128, 27, False),
    "small_units": (300, 64, 3, 128, 64, 0, True),
    "one_expert_tail_rows": (129, 40, 1, 256, 64, 5, True),
    "many_tiles": (128 * 150 + 17, 128, 2, 128, 128, 9, True),
    "wide_gates": (260, 256, 4, 256, 128, 100, False),
}


@pytest.mark.parametrize("name", sorted(CHAIN_CASES))
def test_ple_chain_fwd(name):
    """cdcmdr_ple_chain_fwd (PLE level 0: expert layers 0 -> 1 chained through shared memory + gate logits in one tcgen05 kernel)
    against the host restatement on identical bf16 operands: layer-0 activation and expert outputs within one bf16 ulp of the
    tensor scale, logits 3e-5."""
    B, K0, nE, d0, d1, ng, keep_a0 = CHAIN_CASES[name]

    def fn(lib, e):
        b16 = lambda a: e.put(f32_to_bf16(a.astype(np.float32)).reshape(a.shape).view(np.int16))   # noqa: E731
        ldx = K0 + 8
        X = b16(e.rng.standard_normal((B, ldx)))
        W0 = b16(e.rng.standard_normal((nE * d0 + ng, K0)) / np.sqrt(K0))
        W1 = b16(e.rng.standard_normal((nE * d1, d0)) / np.sqrt(d0))
        b0, b1 = e.f32(nE * d0 + ng, scale=0.3), e.f32(nE * d1, scale=0.3)
        A0 = e.zeros(B, nE * d0, dtype=torch.int16)
        H = e.zeros(B, nE * d1 + 8, dtype=torch.int16)
        Lg = e.zeros(B, max(ng, 1) + 3)
        d = L.PleChain(X.data_ptr(), ldx, B, K0, W0.data_ptr(), b0.data_ptr(), W1.data_ptr(), b1.data_ptr(), nE, d0, d1, ng,
                       A0.data_ptr() if keep_a0 else None, nE * d0, H.data_ptr(), nE * d1 + 8, Lg.data_ptr() if ng else None, max(ng, 1) + 3,
                       0.0, None, 0, 0)
        lib.ple_chain_fwd(C.byref(d), 0)
        return [A0, H, Lg]
    cpu, gpu = Env(31).run(fn)
    for i, what in ((0, "A0"), (1, "H")):
        a = (cpu[i].view(np.uint16).astype(np.uint32) << 16).view(np.float32)
        b = (gpu[i].view(np.uint16).astype(np.uint32) << 16).view(np.float32)
        if what == "A0" and not keep_a0:
            assert not b.any()                                     # inference: the layer-0 activation is never written
            continue
        scale = max(float(np.abs(a).max()), 1e-30)
        assert float(np.abs(a - b).max()) <= 2 ** -7 * scale, (name, what, float(np.abs(a - b).max()), scale)
        assert float((np.abs(a - b) > 2 ** -8 * scale).mean()) < 0.02, (name, what)
    if ng:
        check(cpu[2], gpu[2], tol=3e-5, what=f"{name} logits")


def test_ple_chain_dropout_statistics():
    """nn.Dropout on both chained layers: with X = 0 and positive biases every kept entry equals bias/(1-p); keep rates, the
    1/(1-p) scale (layer 1 sees the dropped layer-0 tile), bit-repeatability and salt dependence."""
    lib = cm._lib.load()
    dev = torch.device("cuda")
    B, K0, nE, d0, d1, p = 2048, 64, 2, 256, 128, 0.25
    X = torch.zeros(B, K0, dtype=torch.bfloat16, device=dev)
    W0 = torch.zeros(nE * d0, K0, dtype=torch.bfloat16, device=dev)
    W1 = torch.zeros(nE * d1, d0, dtype=torch.bfloat16, device=dev)
    b0, b1 = torch.ones(nE * d0, device=dev), torch.ones(nE * d1, device=dev)
    st = torch.zeros(48, dtype=torch.uint8, device=dev)
    lib.step_state_init(st.data_ptr(), 3, 0)
    lib.step_tick(st.data_ptr(), 1e-3, 0.9, 0.99, 1e-8, 0.0, 2000, 0)

    def run(salt):
        A0 = torch.zeros(B, nE * d0, dtype=torch.bfloat16, device=dev)
        H = torch.zeros(B, nE * d1, dtype=torch.bfloat16, device=dev)
        d = L.PleChain(X.data_ptr(), K0, B, K0, W0.data_ptr(), b0.data_ptr(), W1.data_ptr(), b1.data_ptr(), nE, d0, d1, 0,
                       A0.data_ptr(), nE * d0, H.data_ptr(), nE * d1, None, 0, p, st.data_ptr() + 8, salt, salt + 1)
        lib.ple_chain_fwd(C.byref(d), 0)
        torch.cuda.synchronize()
        return A0.float().cpu().numpy(), H.float().cpu().numpy()
    (a, h), (a2, h2), (a3, h3) = run(5), run(5), run(9)
    assert np.array_equal(a, a2) and np.array_equal(h, h2)
    for t in (a, h):
        vals = set(np.unique(t).tolist())
        assert vals == {0.0, float(torch.tensor(1 / (1 - p)).bfloat16())}, vals
        assert abs((t > 0).mean() - (1 - p)) < 4e-3
        k = (t > 0).astype(np.float64) - (t > 0).mean()
        assert abs((k[:, 1:] * k[:, :-1]).mean()) < 2e-3 and abs((k[1:] * k[:-1]).mean()) < 2e-3
    assert 0.30 < ((a > 0) != (a3 > 0)).mean() < 0.45              # independent masks differ on 2p(1-p) = 37.5 %
