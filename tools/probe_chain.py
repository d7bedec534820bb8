"""Timing probe of cdcmdr_ple_chain_fwd at the C4 shape (B = 65 536, K0 = 368, 10 experts 256 -> 128, 27 gate columns) with CUDA
events: training form (layer-0 activation stored, dropout 0.2), inference form (no store, no dropout), and the per-layer launches
it replaces (concatenated-N layer-0 GEMM + grouped layer-1 GEMM + gate GEMM).  One JSON line per variant."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cdcmdr_b200 as cm  # noqa: E402

L = cm._lib
lib = L.load()
dev = torch.device("cuda")
B, K0, nE, d0, d1, ng = int(os.environ.get("PROBE_B", 65536)), 368, 10, 256, 128, 27
torch.manual_seed(0)
X = torch.randn(B, K0 + 8, device=dev).bfloat16()
W0 = (torch.randn(nE * d0 + ng, K0, device=dev) / K0 ** 0.5).bfloat16()
W1 = (torch.randn(nE * d1, d0, device=dev) / d0 ** 0.5).bfloat16()
b0, b1 = torch.randn(nE * d0 + ng, device=dev) * 0.1, torch.randn(nE * d1, device=dev) * 0.1
A0 = torch.empty(B, nE * d0, dtype=torch.bfloat16, device=dev)
H = torch.empty(B, nE * d1, dtype=torch.bfloat16, device=dev)
Lg = torch.empty(B, 32, device=dev)
st = torch.zeros(48, dtype=torch.uint8, device=dev)
lib.step_state_init(st.data_ptr(), 3, 0)
lib.step_tick(st.data_ptr(), 1e-3, 0.9, 0.99, 1e-8, 0.0, 2000, 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()                                   # inputs > L2 anyway; keeps the weights from staying hot between launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def chain(train):
    d = L.PleChain(X.data_ptr(), K0 + 8, B, K0, W0.data_ptr(), b0.data_ptr(), W1.data_ptr(), b1.data_ptr(), nE, d0, d1, ng,
                   A0.data_ptr() if train else None, nE * d0, H.data_ptr(), nE * d1, Lg.data_ptr(), 32,
                   0.2 if train else 0.0, st.data_ptr() + 8 if train else None, 1, 2)
    lib.ple_chain_fwd(C.byref(d), torch.cuda.current_stream().cuda_stream)


flops = 2.0 * B * (K0 * (nE * d0 + ng) + nE * d0 * d1)
for name, train in (("chain_train", True), ("chain_inference", False)):
    us = timed(lambda: chain(train))
    byt = B * (K0 * 2 + nE * d1 * 2 + ng * 4 + (nE * d0 * 2 if train else 0))
    print(json.dumps(dict(name=name, B=B, us=round(us, 1), tflops=round(flops / us / 1e6, 1), hbm_gbs=round(byt / us / 1e3, 1))), flush=True)
# wait counters of the training form (fractions of the CTA lifetime)
ctr = torch.zeros(16, dtype=torch.int64, device=dev)
lib.ple_chain_profile(ctr.data_ptr())
chain(True)
torch.cuda.synchronize()
lib.ple_chain_profile(None)
c = ctr.cpu().tolist()
n = max(c[7], 1)
names = ["producer_wait_slot", "mma_wait_accum", "mma_wait_weights", "mma_wait_act", "epi_wait_accum", "epi_wait_actblock"]
life = dict(epi=c[6] / n, mma=c[8] / n, producer=c[9] / n)
print(json.dumps(dict(name="chain_train_waits", cycles_per_cta=life, frac={k: round(c[i] / n / max(life["epi"], 1), 3) for i, k in enumerate(names)})))
