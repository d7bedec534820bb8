"""GPU timing of the C4 step with the reference's STOCK attention settings switched on (config.py:24-28: use_atten, 64-d, 2 heads,
3 layers, V_res residual) - not a test, not the bench contract:

    python tools/bench_atten.py [--batch 65536] [--steps 5]

CDC(base=PLE) as in bench.py, bf16 tensor-core path for experts / gates / towers AND the attention block (bf16 token matrices,
projections / input gradients / weight gradients on the tcgen05 GEMM, bf16-I/O attention core; CDCMDR_PRECISION=fp32 for the
exact path).  Times the fused training step with and without the block (CUDA events, 2 warm-ups) and the
block's own forward / backward launches.  Prints one JSON line."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cdcmdr_b200 as cm  # noqa: E402

F, E, T, ND, DOM = 23, 16, 4, 30, 10
DEV = os.environ.get("CDCMDR_DEVICE", "cuda")
PREC = os.environ.get("CDCMDR_PRECISION", "bf16")
DROP = float(os.environ.get("CDCMDR_DROPOUT", "0.2"))


def cfg(atten):
    class Cfg:
        use_atten = atten; use_dcn = False; ple_n_expert_specific = 2; ple_n_expert_shared = 2
        atten_embed_dim = 64; att_layer_num = 3; att_head_num = 2; att_res = True
        cdcmdr_precision = PREC
    return Cfg()


def timed(fn, n):
    if DEV != "cuda":
        import time
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        return 1e3 * (time.perf_counter() - t0) / n
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    B = a.batch
    fd = np.full(F, 45_454, dtype=np.int64); fd[DOM] = ND
    rng = np.random.default_rng(3)
    x = np.stack([np.minimum(rng.zipf(1.05, size=B) - 1, c - 1) for c in fd], axis=1).astype(np.int32)
    x[:, DOM] = 7
    xt = torch.from_numpy(x).to(DEV)
    yt = torch.from_numpy((rng.random(B) < 0.05).astype(np.int16)).to(DEV)
    out = dict(what=f"C4 fused step with the stock attention block ({PREC} path)", batch=B)
    for atten in (False, True):
        torch.manual_seed(2000)
        m = cm.CDC(fd, E, T, ND, "ple", ((256, 128), (64,)), (64, 32), DOM, dropout=DROP, config=cfg(atten), l2_reg_embedding=1e-5,
                   l2_reg_linear=1e-5, l2_reg_dnn=1e-5).to(DEV).train()
        m.set_groups([d % T for d in range(ND)])
        opt = cm.Adam(m.parameters(), lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)
        step = lambda: m.train_step(xt, yt, opt, mode="split", domain_i=7)   # noqa: E731
        step(); step()
        ms = timed(step, a.steps)
        key = "with_atten" if atten else "without"
        out[key] = dict(ms_per_step=ms, samples_per_s=B / ms * 1e3)
        if atten:
            base = m.base_model_instance
            rt, att = base._rt, base._att
            ws = rt.ws(B)
            X32 = base._att_x(ws, base._x_mat(ws, B), B)          # bf16 tokens on the tensor-core path, fp32 embeddings otherwise
            lin = ws.mat("bench.lin", B, 1)
            dX = ws.mat("bench.dX", B, F * E)
            out[key]["block_fwd_ms"] = timed(lambda: att.fwd(ws, X32, B, lin, True), a.steps)
            out[key]["block_bwd_ms"] = timed(lambda: att.bwd(ws, X32, B, lin, dX, True), a.steps)
            flop = 2 * B * (F * E * 64 * 2 + 3 * (F * 64 * 192 + 2 * 2 * F * F * 32 + F * 64 * 64) + F * 64)
            out[key]["block_fwd_tflops"] = flop / out[key]["block_fwd_ms"] / 1e9
        del m, opt
        if DEV == "cuda":
            torch.cuda.empty_cache()
    out["mem_gb"] = torch.cuda.max_memory_allocated() / 2 ** 30 if DEV == "cuda" else 0.0
    print(json.dumps(out))


if __name__ == "__main__":
    main()
