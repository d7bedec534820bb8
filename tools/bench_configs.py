"""GPU timing of the BASELINE.json configs other than the bench.py headline (C4): one JSON line per case (not a test, not the
bench contract).  Each case is the fused training step (gather -> model -> BCE -> backward -> regulariser -> Adam, dense-exact
embedding update) recorded into one CUDA graph, bf16 tensor-core path, synthetic Zipf(1.05) ids, dropout 0.2; 3 warm-up + 10
timed replays with CUDA events over 4 rotating batches.

    python tools/bench_configs.py [c1|c2mix|c2v2|c3ple|c3mmoe ...]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cdcmdr_b200 as cm  # noqa: E402

L2 = dict(l2_reg_embedding=1e-5, l2_reg_linear=1e-5, l2_reg_dnn=1e-5)
ADAM = dict(lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)
DROPOUT = 0.2


def cfg(precision):
    class Cfg:
        use_atten = False; use_dcn = False; ple_n_expert_specific = 2; ple_n_expert_shared = 2; mmoe_n_expert = 8
        cdcmdr_precision = precision
    return Cfg()


def batches(fd, B, n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        x = np.stack([np.minimum(rng.zipf(1.05, size=B) - 1, d - 1) for d in fd], axis=1).astype(np.int32)
        y = (rng.random(B) < 0.05).astype(np.int16)
        out.append((torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()))
    return out


def case(name):
    if name == "c1":
        F, E, T, B = 16, 16, 4, 2048
        fd = np.full(F, 66_666, dtype=np.int64); fd[-1] = 4
        build = lambda p: cm.PLE(fd, E, T, 2, 2, ((256, 128), (64,)), (64, 32), dropout=DROPOUT, config=cfg(p), **L2)   # noqa: E731
        kw, prec, what = dict(mode="gather"), "fp32", "C1: PLE fp32, 4 domains, 16 fields x embed 16, vocab 1M, batch 2048"
        sel = lambda x: x[:, -1].long()   # noqa: E731
    elif name in ("c2mix", "c2v2"):
        F, E, B = 26, 32, 16384
        fd = np.full(F, 40_000, dtype=np.int64)
        mix = name == "c2mix"
        build = lambda p: cm.DCNv2(fd, E, 3, (512, 256, 128), dropout=DROPOUT, use_low_rank_mixture=mix, config=cfg(p), **L2)   # noqa: E731
        kw, prec = dict(mode="col", col=0), "bf16"
        what = f"C2: DCNv2 ({'CrossNetMix, stock' if mix else 'CrossNetV2'}), 3 cross layers + MLP 512-256-128, 26 fields x embed 32, batch 16384"
        sel = None
    elif name in ("c3ple", "c3mmoe"):
        F, E, T, B = 23, 16, 3, 65536
        fd = np.full(F, 45_000, dtype=np.int64); fd[10] = 10
        if name == "c3ple":
            build = lambda p: cm.PLE(fd, E, T, 2, 2, ((256, 128), (64,)), (64, 32), dropout=DROPOUT, config=cfg(p), **L2)   # noqa: E731
            what = "C3: PLE 3 tasks x 2 specific + 2 shared (8 experts) x 2 levels, 10 domains, batch 65536"
        else:
            build = lambda p: cm.MMoE(fd, E, T, 8, (256, 128, 64), (64, 32), dropout=DROPOUT, config=cfg(p), **L2)   # noqa: E731
            what = "C3: MMoE 8 experts (256,128,64), 3 tasks, 10 domains, batch 65536"
        kw, prec = dict(mode="gather"), "bf16"
        sel = lambda x: (x[:, 10] % T).long()   # noqa: E731
    else:
        raise SystemExit(f"unknown case {name}")
    torch.manual_seed(2000)
    model = build(prec).to("cuda").train()
    opt = cm.Adam(model.parameters(), **ADAM)
    data = batches(fd, B, 4, 5)
    step = cm.GraphedTrainStep(model, opt, B, len(fd), **kw)
    sel_buf = None
    if sel is not None:
        sel_buf = sel(data[0][0]).contiguous()
        step.kw["sel"] = sel_buf
    step.x.copy_(data[0][0]); step.y.copy_(data[0][1])
    step.capture()

    def run(n):
        for i in range(n):
            x, y = data[i % len(data)]
            step.x.copy_(x); step.y.copy_(y)
            if sel_buf is not None:
                sel_buf.copy_(sel(x))
            step()
    run(3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record(); run(n); e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    loss = model.step_losses(step.out)[0]
    print(json.dumps(dict(workload=what, dtype=prec, batch=B, ms_per_step=round(ms, 4), samples_per_s=round(B / ms * 1e3),
                          launches_per_step=step.launches_per_step, loss_last=round(float(loss), 5), dropout=DROPOUT,
                          embedding_update="dense_exact", data="synthetic Zipf(1.05)")), flush=True)


if __name__ == "__main__":
    for nm in (sys.argv[1:] or ["c1", "c2mix", "c2v2", "c3ple", "c3mmoe"]):
        case(nm)
