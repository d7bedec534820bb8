"""BASELINE configs[4] ("C5"): STAR / CDC(base=STAR) with a ROW-RANGE sharded embedding table (default 500 M rows x embed 64 over
the ranks: 16 GB of fp32 rows + 32 GB of Adam moments per GPU at N = 8), towers (256,128,64,32), 23 fields, global batch 65 536,
bf16 tensor-core path, dropout 0.2, touched-row Adam with the incrementally maintained regulariser (the reference-exact dense
update would sweep 96 GB per GPU per step).  Not the bench.py contract - a SCALE-style measurement of its own:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_c5.py [--rows 500000000]
                                    [--global-batch 65536] [--model cdc|star] [--steps 20] [--gather-only]

Every rank allocates ONLY its rows (parallel.sharded_table); the forward is one peer-memory gather kernel; the backward
all-gathers indices + bf16 row gradients.  Prints one JSON line (rank 0): samples/s over the global batch (CUDA events, max over
ranks), per-phase times of one eager step, the gather kernel's GB/s with uniform and Zipf ids, memory per rank."""
import argparse
import datetime
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cdcmdr_b200 as cm  # noqa: E402

F, E, T, ND, DOM = 23, 64, 4, 30, 10
TOWER = (256, 128, 64, 32)
L2 = dict(l2_reg_embedding=1e-5, l2_reg_linear=1e-5, l2_reg_dnn=1e-5)
ADAM = dict(lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)


class Cfg:
    use_atten = False; use_dcn = False; cdcmdr_precision = "bf16"


def timed(fn, n, dev, world):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    return ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=500_000_000)
    ap.add_argument("--global-batch", type=int, default=65536)
    ap.add_argument("--model", default="cdc", choices=["cdc", "star"])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--gather-only", action="store_true")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if not dist.is_initialized():
        if "MASTER_ADDR" not in os.environ:
            os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev, timeout=datetime.timedelta(seconds=300))
    fd = np.full(F, (a.rows - ND) // (F - 1), dtype=np.int64)
    fd[DOM] = ND
    B = a.global_batch // world
    torch.manual_seed(2000 + rank)
    with cm.parallel.sharded_table(rank, world, dev):
        if a.model == "cdc":
            model = cm.CDC(fd, E, T, ND, "star", None, TOWER, DOM, dropout=0.2, config=Cfg(), **L2)
        else:
            model = cm.STAR(fd, E, T, TOWER, domain_idx=DOM, dropout=0.2, config=Cfg(), **L2)
    model = model.to(dev).train()
    base = getattr(model, "base_model_instance", model)
    if a.model == "cdc":
        model.set_groups([d % T for d in range(ND)])
    dp = cm.parallel.attach_data_parallel(model, shard_embedding="rows")
    opt = cm.Adam(model.parameters(), **ADAM)
    rng = np.random.default_rng([7, rank])
    nb = 4

    def batch(zipf=True):
        cols = [np.minimum(rng.zipf(1.05, size=B) - 1, d - 1) if zipf else rng.integers(0, d, size=B) for d in fd]
        x = np.stack(cols, axis=1).astype(np.int32)
        x[:, DOM] = 7
        y = (rng.random(B) < 0.05).astype(np.int16)
        return torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
    batches = [batch() for _ in range(nb)]
    out = dict(what=f"C5: {'CDC(base=STAR)' if a.model == 'cdc' else 'STAR'}, row-range sharded table in NVLink peer memory",
               n_gpus=world, rows_total=int(fd.sum()), rows_per_gpu=int(base.embedding.rows_per), embed_dim=E, fields=F,
               global_batch=B * world, batch_per_gpu=B, table_gb_per_gpu=round(base.embedding.rows_per * E * 4 / 2 ** 30, 2),
               dtype="bf16", dropout=0.2, embedding_update="sparse_lazy (touched rows; labelled deviation from the reference's dense Adam)")
    # ---- the gather kernel alone (rows where they live; table >> L2): uniform ids = worst case, Zipf = the workload's skew
    rt = base._rt
    ws = rt.ws(B)
    X = base._x_mat(ws, B)
    for tag, zipf in (("uniform", False), ("zipf", True)):
        xg, _ = batch(zipf)
        fn = lambda: dp.embed_forward(ws, xg, B, X, plan_ahead=False)   # noqa: E731
        for _ in range(3):
            fn()
        ms = timed(fn, 20, dev, world)
        out[f"gather_{tag}"] = dict(us=round(ms * 1e3, 1), gbs_per_gpu=round(B * F * (4 + E * 4 + E * 2) / ms / 1e6, 1),
                                    algorithmic_bytes=B * F * (4 + E * 4 + E * 2))
    if not a.gather_only:
        kw = dict(mode="split", domain_i=7) if a.model == "cdc" else dict(mode="col", col=0)
        step = cm.GraphedTrainStep(model, opt, B, F, **kw)
        step.x.copy_(batches[0][0]); step.y.copy_(batches[0][1])
        step.capture()
        it = [0]

        def run():
            x, y = batches[it[0] % nb]
            it[0] += 1
            step.x.copy_(x); step.y.copy_(y)
            step()
        for _ in range(a.warmup):
            run()
        ms = timed(run, a.steps, dev, world)
        loss = base.step_losses(step.out)
        out.update(ms_per_step=round(ms, 4), samples_per_s=round(B * world / ms * 1e3), launches_per_step=step.launches_per_step,
                   loss_last=round(float(loss[0]), 5), bce_last=round(float(loss[1]), 5))
    out["mem_gb_per_gpu"] = round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
