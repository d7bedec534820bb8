"""GPU probe: times cdcmdr_gemm_bf16_tc at the C4 shapes with epilogue variants (not a test, not the bench)."""
import ctypes as C
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cdcmdr_b200 as cm

L = cm._lib
lib = L.load()
dev = torch.device("cuda")
st = torch.zeros(48, dtype=torch.uint8, device=dev)
lib.step_state_init(st.data_ptr(), 1, 0)
seed_ptr = st.data_ptr() + 8


def run(name, M, N, K, a_mn=0, b_mn=0, out="bf16", act=0, drop=0.0, bias=True, mask=False, split=1, block_n=0, reps=10, pad=0, mode=None):
    """pad: round the operands' row pitch up to a multiple of `pad` elements (logical shape unchanged)"""
    def pitch(n):
        return (n + pad - 1) // pad * pad if pad else n
    if mode is not None:
        lib.gemm_bf16_tc_mode(mode)
    if a_mn:
        A = torch.randn(K, pitch(M), device=dev).to(torch.bfloat16); lda, ar, ac = pitch(M), K, M
    else:
        A = torch.randn(M, pitch(K), device=dev).to(torch.bfloat16); lda, ar, ac = pitch(K), M, K
    if b_mn:
        B = torch.randn(K, pitch(N), device=dev).to(torch.bfloat16); ldb, br, bc = pitch(N), K, N
    else:
        B = torch.randn(N, pitch(K), device=dev).to(torch.bfloat16); ldb, br, bc = pitch(K), N, K
    bias_t = torch.randn(N, device=dev) if bias else None
    n_main = N if out == "bf16" else 0
    om = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if n_main else None
    s = lib.gemm_bf16_tc_splits(K, split) if split > 1 else 1
    oa = torch.empty(max(s, 1) * M * N, device=dev) if not n_main else None
    mk = torch.randn(M, N, device=dev).to(torch.bfloat16) if mask else None
    d = L.GemmBf16(A.data_ptr(), lda, ar, ac, B.data_ptr(), ldb, br, bc, M, N, K, 1, 0, 0, 0, 0, a_mn, b_mn,
                   bias_t.data_ptr() if bias else None, 0, n_main, om.data_ptr() if om is not None else None, N, 0,
                   oa.data_ptr() if oa is not None else None, N, 0, act, mk.data_ptr() if mask else None, N, 0, 1.25,
                   drop, seed_ptr if drop > 0 else None, 3, 0, s, M * N, block_n)
    for _ in range(3):
        lib.gemm_bf16_tc(C.byref(d), 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        lib.gemm_bf16_tc(C.byref(d), 0)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    prof = None
    if os.environ.get("PROBE_PROF"):
        cnt = torch.zeros(16, dtype=torch.int64, device=dev)
        lib.gemm_bf16_tc_profile(cnt.data_ptr())
        lib.gemm_bf16_tc(C.byref(d), 0)
        torch.cuda.synchronize()
        lib.gemm_bf16_tc_profile(None)
        c = cnt.tolist()
        n = max(c[7], 1)
        prof = dict(cta_kcyc=round(c[6] / n / 1e3, 1), prod_wait_empty=round(c[0] / c[6], 3), mma_wait_tempty=round(c[1] / c[6], 3),
                    mma_wait_full=round(c[2] / c[6], 3), epi_wait_tfull=round(c[3] / c[6], 3), epi_wait_stage=round(c[4] / c[6], 3),
                    epi_life=round(c[5] / c[6], 3), epi_tmem_ld=round(c[8] / c[6], 3), epi_math=round(c[9] / c[6], 3),
                    epi_stage_store=round(c[10] / c[6], 3), epi_bias=round(c[11] / c[6], 3))
    tf = 2.0 * M * N * K / (us * 1e-6) / 1e12
    by = (M * K + N * K) * 2 + M * N * (2 if n_main else 4 * s) + (M * N * 2 if mask else 0)
    print(json.dumps(dict(name=name, M=M, N=N, K=K, us=round(us, 1), tflops=round(tf, 1), gbs=round(by / us / 1e3, 1), pad=pad, mode=mode, prof=prof)), flush=True)
    if mode is not None:
        lib.gemm_bf16_tc_mode(0)


B = 65536
ONLY = sys.argv[1] if len(sys.argv) > 1 else None
if ONLY:
    _run = run

    def run(name, *a, **k):            # noqa: F811
        if name.startswith(ONLY):
            _run(name, *a, **k)
# epilogue knock-out probes (mode bits 16 / 64 / 32: no drain / tcgen05.ld only / everything but the TMA store)
for dbg in (0, 16, 64, 32):
    run(f"dbg{dbg}_l0_fwd_full", B, 2560, 368, act=1, drop=0.2, mode=1 | dbg)
    run(f"dbg{dbg}_l0_fwd_full_pad", B, 2560, 368, act=1, drop=0.2, pad=64, mode=1 | dbg)
    run(f"dbg{dbg}_l0_fwd_pair", B, 2560, 368, act=1, drop=0.2, mode=4 | dbg)
for md in (1, 4):
    for pd in (0, 64):
        run("align_l0_fwd_plain", B, 2560, 368, bias=False, pad=pd, mode=md)
        run("align_l0_fwd_full", B, 2560, 368, act=1, drop=0.2, pad=pd, mode=md)
        run("align_l0_dgrad", B, 368, 2560, b_mn=1, out="f32", pad=pd, mode=md)
        run("align_l0_wgrad", 2560, 368, B, a_mn=1, b_mn=1, out="f32", split=6, pad=pd, mode=md)
run("l0_fwd_plain", B, 2560, 368, bias=False)
run("l0_fwd_bias", B, 2560, 368)
run("l0_fwd_bias_relu", B, 2560, 368, act=1)
run("l0_fwd_bias_relu_drop", B, 2560, 368, act=1, drop=0.2)
run("l0_fwd_bn128", B, 2560, 368, act=1, block_n=128)
run("l0_fwd_f32out", B, 2560, 368, out="f32")
run("l0b_fwd(256->128 x10 as one)", B, 128, 256, act=1, drop=0.2)
run("l0_dgrad (N=368,K=2560,b_mn)", B, 368, 2560, b_mn=1, out="f32")
run("l0_dgrad_bf16", B, 368, 2560, b_mn=1)
run("l0_wgrad (M=2560,N=368,K=B)", 2560, 368, B, a_mn=1, b_mn=1, out="f32", split=6)
run("l0_wgrad_nosplit", 2560, 368, B, a_mn=1, b_mn=1, out="f32", split=1)
run("l1_mask_dgrad (N=256,K=128)", B, 256, 128, b_mn=1, mask=True)
run("big_square", 8192, 8192, 8192, bias=False)


def colsum_probe(Bn, Cn, bf16=True):
    X = torch.randn(Bn, Cn, device=dev)
    if bf16:
        X = X.to(torch.bfloat16)
    out = torch.zeros(Cn, device=dev)
    sc = torch.empty(lib.colsum_scratch_bytes(Cn), dtype=torch.uint8, device=dev)
    for _ in range(3):
        lib.colsum(X.data_ptr(), Cn, 1 if bf16 else 0, Bn, Cn, out.data_ptr(), 0, sc.data_ptr(), 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        lib.colsum(X.data_ptr(), Cn, 1 if bf16 else 0, Bn, Cn, out.data_ptr(), 0, sc.data_ptr(), 0)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    ref = X.float().sum(0)
    err = float((out - ref).abs().max() / ref.abs().max())
    print(json.dumps(dict(name=f"colsum {Bn}x{Cn} {'bf16' if bf16 else 'f32'}", us=round(us, 1), gbs=round(X.numel() * X.element_size() / us / 1e3, 1), relerr=err)), flush=True)


if not ONLY:
    colsum_probe(B, 2560)
    colsum_probe(B, 1280)
    colsum_probe(B, 640, bf16=False)
