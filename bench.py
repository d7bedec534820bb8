#!/usr/bin/env python
"""bench.py - CDC-PLE training throughput (fwd + bwd + optimizer) on synthetic Ali-CCP-shaped multi-domain data.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|bf16]

One JSON line on stdout (rank 0).  Workload = BASELINE.json configs[3] ("C4"): CDC over 30 domains on a PLE backbone
(4 clusters, expert dims ((256,128),(64,)), towers (64,32)), 23 sparse fields x embed 16, vocab 1M, batch 65,536 per
GPU, steady-state step `model(X, mode='split', domain_i=d)` + BCE + L2 + Adam (reference run.py:635-640).
`value` times K steps with inputs resident in HBM; `e2e` times the same step through the public API from pinned HOST
buffers (H2D of the batch and D2H of the loss inside the timed region, one synchronisation per step like the reference's
loss.item()).  `--impl reference` times the reference's OWN torch modules (oracle/_ref, installed verbatim by oracle/make_ref.py) on the host
cores at the same batch size; the numpy oracle port is reported beside it (and is the fallback when oracle/_ref is absent).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F, E, T, N_DOMAIN, DOMAIN_IDX = 23, 16, 4, 30, 10
VOCAB_TOTAL = 1_000_000
EXPERT_DIMS, TOWER_DIMS = ((256, 128), (64,)), (64, 32)
NS, NSH = 2, 2
L2 = dict(l2_reg_embedding=1e-5, l2_reg_linear=1e-5, l2_reg_dnn=1e-5)
ADAM = dict(lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)
DROPOUT = 0.2                       # reference default (model/ple.py:17); active in train mode
SEED = 2000


def field_dims():
    fd = np.full(F, (VOCAB_TOTAL - N_DOMAIN) // (F - 1), dtype=np.int64)
    fd[DOMAIN_IDX] = N_DOMAIN
    return fd


def make_batches(n_batches, B, seed, rank=0):
    """Ali-CCP-shaped: Zipf(1.05) ids per field, each batch holds ONE domain (run.py:499-526 per-domain loaders),
    domains drawn from a power law, labels Bernoulli(0.05).  Data-parallel ranks hold different rows of the SAME step's
    domain batch, so the domain sequence comes from the shared seed and everything else from (seed, rank)."""
    drng = np.random.default_rng(seed)
    rng = np.random.default_rng([seed, rank])
    fd = field_dims()
    pw = 1.0 / np.arange(1, N_DOMAIN + 1) ** 1.2
    pw /= pw.sum()
    out = []
    for _ in range(n_batches):
        x = np.empty((B, F), dtype=np.int32)
        for f in range(F):
            if f == DOMAIN_IDX:
                continue
            x[:, f] = np.minimum(rng.zipf(1.05, size=B) - 1, fd[f] - 1).astype(np.int32)
        d = int(drng.choice(N_DOMAIN, p=pw))
        x[:, DOMAIN_IDX] = d
        y = (rng.random(B) < 0.05).astype(np.int16)
        out.append((x, y, d))
    return out


class Cfg:
    use_atten = False
    use_dcn = False
    ple_n_expert_specific = NS
    ple_n_expert_shared = NSH
    cdcmdr_precision = "fp32"


def fwd_flops_per_sample():
    D = F * E
    nE = T * NS + NSH
    w = NS + NSH
    f = D * (nE * 256 + T * w + nE + 1) + nE * 256 * 128                    # level 0 experts + gates + wide linear
    f += T * 128 * (NS * 64 + w) + 128 * NSH * 64                           # level 1
    f += T * (64 * 64 + 64 * 32 + 32)                                       # towers
    return 2 * f


class ClockSampler:
    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def mark(self):
        return len(self.rows)

    def stop(self, lo=0):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows[lo:]:
            p = [c.strip() for c in r.split(",")]
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except Exception:
                continue
            for n, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))


# ----------------------------------------------------------------------------------------------- CPU arm
def run_cpu_port(steps, warmup, B_sample):
    """The reference's algorithm for the path restated in numpy (oracle/, pinned against reference-generated golden
    fixtures), with the reference's default dropout, on all host cores numpy's BLAS uses."""
    from oracle import cdcmdr_oracle as O
    rng = np.random.default_rng(SEED)
    fd = field_dims()
    om = O.PLE(fd, E, T, NS, NSH, EXPERT_DIMS, TOWER_DIMS, **L2)
    keep = 1.0 - DROPOUT
    om.drop = lambda shape: (rng.random(shape, dtype=np.float32) < keep).astype(np.float32) / np.float32(keep)
    sd = init_state_numpy(fd, rng)
    opt = O.Adam(**{k: ADAM[k] for k in ("lr", "betas", "eps", "weight_decay")})
    batches = make_batches(max(1, min(4, steps + warmup)), B_sample, SEED + 1)
    d2g = np.arange(N_DOMAIN) % T
    t_steps = []
    for i in range(warmup + steps):
        x, y, d = batches[i % len(batches)]
        t0 = time.perf_counter()
        O.train_step(om, sd, opt, x, y, "split_domain", domain2group=d2g, domain_i=d)
        if i >= warmup:
            t_steps.append(time.perf_counter() - t0)
    sec = float(np.sum(t_steps))
    return B_sample * steps / sec, sec / steps


def init_state_numpy(fd, rng):
    """Random-init weights of the reference architecture with the reference's state_dict names."""
    D, sd = F * E, {}

    def lin(name, out, inp):
        b = 1.0 / np.sqrt(inp)
        sd[name + ".weight"] = rng.uniform(-b, b, size=(out, inp)).astype(np.float32)
        sd[name + ".bias"] = rng.uniform(-b, b, size=(out,)).astype(np.float32)

    sd["embedding.embedding_dict.weight"] = rng.standard_normal((int(fd.sum()), E)).astype(np.float32)
    lin("linear.fc", 1, D)
    nE, w = T * NS + NSH, NS + NSH
    for l, dims in enumerate(EXPERT_DIMS):
        inp0 = D if l == 0 else EXPERT_DIMS[l - 1][-1]
        for kind, cnt in (("experts_specific", T * NS), ("experts_shared", NSH)):
            for i in range(cnt):
                inp = inp0
                for j, d in enumerate(dims):
                    lin(f"cgc_layers.{l}.{kind}.{i}.layers.{3 * j}", d, inp)
                    inp = d
        for t in range(T):
            lin(f"cgc_layers.{l}.gates_specific.{t}.0", w, inp0)
        if l + 1 < len(EXPERT_DIMS):
            lin(f"cgc_layers.{l}.gate_shared.0", nE, inp0)
    for t in range(T):
        inp = EXPERT_DIMS[-1][-1]
        for j, d in enumerate(TOWER_DIMS):
            lin(f"towers.{t}.layers.{4 * j}", d, inp)
            k = f"towers.{t}.layers.{4 * j + 1}"
            sd[k + ".weight"] = np.ones(d, np.float32); sd[k + ".bias"] = np.zeros(d, np.float32)
            sd[k + ".running_mean"] = np.zeros(d, np.float32); sd[k + ".running_var"] = np.ones(d, np.float32)
            sd[k + ".num_batches_tracked"] = np.zeros((), np.int64)
            inp = d
        lin(f"towers.{t}.layers.{4 * len(TOWER_DIMS)}", 1, inp)
    return sd


def run_cpu_reference(steps, warmup, B, budget_s=None):
    """The UNMODIFIED reference (its own torch modules, installed verbatim into oracle/_ref/ by oracle/make_ref.py) through its own
    public API on the host cores: CDC(base='ple') with the nested expert dims, loop body exactly run.py:635-640
    (`model(X, mode='split', domain_i=d)`, BCELoss + get_regularization_loss, zero_grad / backward / torch.optim.Adam.step,
    loss.item()), fp32, dropout 0.2, all host threads.  Batches are pre-collated tensors (DataLoader excluded, as SURVEY §8d).
    budget_s: stop timing after this many seconds of timed steps (at least one) - the bounded sample of the cpu_baseline leg.
    Returns (samples/s, seconds per step, timed steps, threads)."""
    import tempfile
    import torch
    from oracle import make_ref
    ref = make_ref.load_reference()
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)

    class RefCfg:
        use_atten = False; use_dcn = False; dataset_name = "synthetic"
        ple_n_expert_specific = NS; ple_n_expert_shared = NSH; mmoe_n_expert = 8
        p_weight = 0.1; p_weight_method = "linear_decay"; p_weight_exp_decay = 0.9; old_matrix_weight = 0.0
        affinity_func = "minus"
    fd = field_dims()
    torch.manual_seed(SEED)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)                                    # CDC.__init__ creates result/<dataset>/ (model/cdc.py:60-62)
        try:
            model = ref["CDC"](fd, E, T, N_DOMAIN, "ple", EXPERT_DIMS, TOWER_DIMS, DOMAIN_IDX,
                               domain_cnt_weight=[1.0 / N_DOMAIN] * N_DOMAIN, dropout=DROPOUT, config=RefCfg(), **L2)
        finally:
            os.chdir(cwd)
    d2g = [d % T for d in range(N_DOMAIN)]
    model.domain2group_list = d2g
    model.domain2group = torch.tensor(d2g, dtype=torch.int64)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=ADAM["lr"], betas=ADAM["betas"], eps=ADAM["eps"], weight_decay=ADAM["weight_decay"])
    crit = torch.nn.BCELoss()
    batches = [(torch.from_numpy(x), torch.from_numpy(y), d) for x, y, d in make_batches(max(1, min(4, steps + warmup)), B, SEED + 1)]
    t_steps = []
    for i in range(warmup + steps):
        x, y, d = batches[i % len(batches)]
        t0 = time.perf_counter()
        pred = model(x, mode="split", domain_i=d)
        loss = crit(pred, y.float())
        loss = loss + model.get_regularization_loss(device="cpu")
        model.zero_grad()
        loss.backward()
        opt.step()
        loss.item()
        if i >= warmup:
            t_steps.append(time.perf_counter() - t0)
            if budget_s is not None and sum(t_steps) >= budget_s:
                break
    sec = float(np.sum(t_steps))
    return B * len(t_steps) / sec, sec / len(t_steps), len(t_steps), threads


def reference_arm(args):
    """`--impl reference`: the reference's own CPU implementation of the path on the box's host cores, same config (batch included)
    as the GPU arm.  Falls back to the numpy port only when oracle/_ref is absent (kind says which)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ.pop("OMP_NUM_THREADS", None)               # torchrun pins it to 1; too late for this process' OpenMP pool, hence set_num_threads below
    from oracle import make_ref
    B = args.batch
    if make_ref.available():
        # a 65 536-row reference step takes 5-15 s of host time: one warm-up step is enough for a CPU loop, and the timed steps stop
        # early once they have used args.ref_budget seconds (the line then reports the steps actually timed)
        warm = min(args.warmup, 1)
        v, sec, done, cores = run_cpu_reference(args.steps, warm, B, budget_s=args.ref_budget)
        kind = "reference"
        sample = (f"{done} of {args.steps} requested steps (after {warm} warm-up; time budget {args.ref_budget:.0f} s) of batch {B}: the "
                  f"reference's own torch modules (oracle/_ref, unmodified) on {cores} threads, fp32, dropout {DROPOUT}, loop body run.py:635-640")
        args.steps, args.warmup = done, warm
    else:
        B = args.cpu_batch
        v, sec = run_cpu_port(args.steps, args.warmup, B)
        kind, cores = "port", os.cpu_count()
        sample = f"{args.steps} steps of batch {B} of the same workload (numpy oracle port, dropout {DROPOUT}); oracle/_ref absent"
    line = {"impl": "reference", "metric": "train samples/sec (fwd+bwd+opt) CDC-PLE", "value": v, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, B),
            "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def workload_config(args, B):
    return {"workload": "C4: CDC(base=PLE) split step, 30 domains -> 4 clusters, 23 fields x embed 16, vocab 1M, "
                        "experts ((256,128),(64,)) x (4x2 specific + 2 shared), towers (64,32)",
            "batch_per_gpu": B, "global_batch": B * args.gpus, "dropout": DROPOUT, "embedding_update": "dense_exact",
            "l2_cache": "working set per step (activations + 64 MB table + Adam moments, > 2 GB) exceeds the 126 MB L2; "
                        "8 distinct batches rotate", "parallelism": f"dp{args.gpus}"}


# ----------------------------------------------------------------------------------------------- GPU arm
def ours_arm(args):
    import torch
    import cdcmdr_b200 as cm

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import datetime
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B = args.batch
    Cfg.cdcmdr_precision = args.precision
    torch.manual_seed(SEED)
    model = cm.CDC(field_dims(), E, T, N_DOMAIN, "ple", EXPERT_DIMS, TOWER_DIMS, DOMAIN_IDX, dropout=DROPOUT, config=Cfg(), **L2)
    model = model.to(dev).train()
    d2g = [d % T for d in range(N_DOMAIN)]
    model.set_groups(d2g)
    base = model.base_model_instance
    lib = base._rt.ops.lib
    if world > 1:
        cm.parallel.attach_data_parallel(base, dist.group.WORLD)
    opt = cm.Adam(model.parameters(), **ADAM)
    nb = 8
    batches = make_batches(nb, B, SEED + 1, rank)
    dev_batches = [(torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev), d) for x, y, d in batches]
    pin_batches = [(torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory(), d) for x, y, d in batches]
    # one CUDA graph per cluster column (the selected tower is a kernel argument), sharing the static input buffers
    # N > 1: the recorded step issues the NEXT batch's embedding exchange behind its own table update (GraphedTrainStep prefetch=True;
    # the loops below hand it batch i+1 with batch i, as a prefetching loader would) - every step still performs exactly one exchange
    pipelined = world > 1 and os.environ.get("CDCMDR_PREFETCH_EXCHANGE", "1") != "0"
    steps_by_col = {}
    proto = None
    for x, y, d in dev_batches:
        col = d2g[d]
        if col in steps_by_col:
            continue
        g = cm.GraphedTrainStep(model, opt, B, F, mode="split", domain_i=d, prefetch=pipelined)
        if proto is not None:
            g.x, g.y, g.x_next = proto.x, proto.y, proto.x_next
        proto = proto or g
        g.x.copy_(x); g.y.copy_(y)
        g.capture()
        steps_by_col[col] = g
    launches_per_step = max(g.launches_per_step for g in steps_by_col.values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    resident_it = [0]

    def run_resident(n):
        for _ in range(n):
            i = resident_it[0]
            x, y, d = dev_batches[i % nb]
            g = steps_by_col[d2g[d]]
            g.x.copy_(x); g.y.copy_(y)
            if pipelined:
                if i == 0:
                    g.prime()                                   # pipeline fill: the first batch's exchange, once
                g.x_next.copy_(dev_batches[(i + 1) % nb][0])
            g()
            resident_it[0] = i + 1

    sums_host = torch.zeros(4, dtype=torch.float64).pin_memory()

    # End to end: every step's batch travels from pinned host memory inside the timed region.  The copy of batch i+1 runs on a
    # copy stream into one of two staging buffers while step i computes (the way a DataLoader with pin_memory + non_blocking
    # copies is meant to be used); the step's static input buffers are then filled device-to-device.
    copy_stream = torch.cuda.Stream()
    stage = [(torch.empty_like(proto.x), torch.empty_like(proto.y)) for _ in range(3 if pipelined else 2)]

    def run_e2e_pipelined(n):
        """The same, one batch deeper: step i needs batch i (its inputs) AND batch i+1's indices (exchanged behind step i's table
        update), so the upload of batch i+2 is what runs while step i computes.  Every batch still crosses H2D inside the timed
        region, every step's loss is read back before the next step is launched."""
        main = torch.cuda.current_stream()

        def upload(i):
            x, y, _ = pin_batches[i % nb]
            sx, sy = stage[i % 3]
            copy_stream.wait_stream(main)                       # the staging buffer's previous readers (step i-3, i-2) are behind us
            with torch.cuda.stream(copy_stream):
                sx.copy_(x, non_blocking=True); sy.copy_(y, non_blocking=True)
                ev = torch.cuda.Event(); ev.record(copy_stream)
            return ev
        evs = {0: upload(0), 1: upload(1)}
        for i in range(n):
            _, _, d = pin_batches[i % nb]
            g = steps_by_col[d2g[d]]
            main.wait_event(evs.pop(i))
            sx, sy = stage[i % 3]
            g.x.copy_(sx); g.y.copy_(sy)
            if i == 0:
                g.prime()
            main.wait_event(evs[i + 1])
            g.x_next.copy_(stage[(i + 1) % 3][0])
            evs[i + 2] = upload(i + 2)
            out = g()
            sums_host.copy_(out["sums"], non_blocking=True)
            main.synchronize()                                  # the reference reads loss.item() every step (run.py:641)

    def run_e2e(n):
        if pipelined:
            return run_e2e_pipelined(n)
        main = torch.cuda.current_stream()

        def upload(i):
            x, y, _ = pin_batches[i % nb]
            sx, sy = stage[i % 2]
            copy_stream.wait_stream(main)                       # the staging buffer's previous reader (step i-2) is behind us
            with torch.cuda.stream(copy_stream):
                sx.copy_(x, non_blocking=True); sy.copy_(y, non_blocking=True)
                ev = torch.cuda.Event(); ev.record(copy_stream)
            return ev
        ev = upload(0)
        for i in range(n):
            _, _, d = pin_batches[i % nb]
            g = steps_by_col[d2g[d]]
            main.wait_event(ev)
            sx, sy = stage[i % 2]
            g.x.copy_(sx); g.y.copy_(sy)
            if i + 1 < n:
                ev = upload(i + 1)
            out = g()
            sums_host.copy_(out["sums"], non_blocking=True)
            main.synchronize()                                  # the reference reads loss.item() every step (run.py:641)

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    run_resident(args.warmup)
    barrier()
    mark = clocks.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_resident(args.steps)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop(mark) if rank == 0 else None
    run_e2e(max(1, args.warmup // 2))
    barrier()
    t0 = time.perf_counter()
    e0.record()
    run_e2e(args.steps)
    e1.record()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), 0.0)
    wall_e2e = (time.perf_counter() - t0) * 1e3
    loss_last = base.step_losses(dict(sums=sums_host, B=B * world, l2_table=base._l2_table()))[0]
    if world > 1:
        t = torch.tensor([ms, max(ms_e2e, wall_e2e)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    else:
        ms_e2e = max(ms_e2e, wall_e2e)
    if args.trace:
        kernel_trace(args.trace, run_resident, barrier, rank, torch)
    roof = dominant_kernel_roofline(base, B, torch) if rank == 0 else None
    emb = embedding_bandwidth(base, dev_batches[0][0], B, torch) if rank == 0 else None
    amort = None
    if world == 1 and not args.no_amortised:
        steps_by_col.clear()                                   # the probing loop regroups: the recorded graphs are stale anyway
        proto = g = None
        amort = cdc_amortised(model, opt, B, ms / args.steps, torch)
    if world > 1:
        # tear-down of a process group while CUDA graphs that recorded its collectives are alive hangs (seen on 2 x B200):
        # drop the graphs, meet once more, and leave without the group destructor
        steps_by_col.clear()
        proto = g = None
        import gc
        gc.collect()
        barrier()
    if rank != 0:
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    value = B * world * args.steps / (ms * 1e-3)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import make_ref
        if make_ref.available():
            # bounded sample: one warm-up + up to 3 timed steps (about 10-30 s) of the SAME batch size on the reference's own modules
            v, sec, done, cores = run_cpu_reference(3, 1, B, budget_s=20.0)
            cpu = {"value": v, "unit": "samples/s", "cores": cores, "kind": "reference",
                   "sample": f"{done} steps (after 1 warm-up) of batch {B} of the same workload on the reference's own torch modules "
                             f"(oracle/_ref, unmodified), fp32, dropout {DROPOUT}; {sec * 1e3:.0f} ms/step"}
            pv, psec = run_cpu_port(1, 1, args.cpu_batch)
            cpu["port"] = {"value": pv, "unit": "samples/s", "sample": f"numpy oracle port, 1 step of batch {args.cpu_batch}"}
        else:
            v, sec = run_cpu_port(2, 1, args.cpu_batch)
            cpu = {"value": v, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": f"2 steps (after 1 warm-up) of batch {args.cpu_batch} of the same workload, numpy oracle port, "
                             f"dropout {DROPOUT}; {sec * 1e3:.0f} ms/step (oracle/_ref absent)"}
    flops_step = 3 * fwd_flops_per_sample() * B
    line = {"metric": "train samples/sec (fwd+bwd+opt) CDC-PLE", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
            "config": workload_config(args, B), "clocks": clk,
            "e2e": {"value": B * world * args.steps / (ms_e2e * 1e-3), "unit": "samples/s",
                    "h2d_bytes_per_step": B * F * 4 + B * 2, "d2h_bytes_per_step": 32, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
            "step_tflops": flops_step / (ms / args.steps * 1e-3) / 1e12, "loss_last": loss_last,
            "roofline": roof, "embedding": emb, "cdc_amortised": amort, "cpu_baseline": cpu, "lib": lib.path}
    if roof is not None and peaks:
        roof["peak_source"] = "MEASURED_PEAKS.json (driver-measured on this pool)"
    if pipelined:
        line["config"]["embedding_exchange"] = ("pipelined: each step issues the NEXT batch's index / row exchange behind its own table "
                                                "update (one exchange per step; CDCMDR_PREFETCH_EXCHANGE=0 runs it at the head of the step)")
    emit(line)
    if world > 1:
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


def kernel_trace(path, run_resident, barrier, rank, torch, steps=4):
    """Diagnostic (not a bench value): CUPTI kernel records of a few resident steps on every rank, rank 0 writes per-kernel totals
    per stream and the busy / idle time of the step's span.  This is how an N-rank step is read when ncu cannot be used on it."""
    from torch.profiler import profile, ProfilerActivity
    barrier()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        run_resident(steps)
        barrier()
    if rank != 0:
        return
    prof.export_chrome_trace(path + ".chrome.json")
    ev = [e for e in prof.events() if e.device_type is not None and "cuda" in str(e.device_type).lower()]
    rows = {}
    spans = []
    for e in ev:
        t0, t1 = e.time_range.start, e.time_range.end
        spans.append((t0, t1))
        k = (e.name[:90])
        r = rows.setdefault(k, [0, 0.0])
        r[0] += 1; r[1] += (t1 - t0)
    spans.sort()
    busy, cur0, cur1 = 0.0, None, None
    for t0, t1 in spans:
        if cur1 is None or t0 > cur1:
            if cur1 is not None:
                busy += cur1 - cur0
            cur0, cur1 = t0, t1
        else:
            cur1 = max(cur1, t1)
    if cur1 is not None:
        busy += cur1 - cur0
    total = spans[-1][1] - spans[0][0] if spans else 0.0
    with open(path, "w") as f:
        f.write(f"# {steps} steps: span {total:.1f} us, union of kernel intervals {busy:.1f} us, sum of kernel times "
                f"{sum(r[1] for r in rows.values()):.1f} us\n")
        for k, (n, t) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{t / steps:10.1f} us/step {n / steps:6.1f} launches/step  {k}\n")


def _time_launch(fn, torch, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / n


def _ncu_traffic(name, shape):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the named kernel from the committed `ncu --set full` capture
    (profiles/r2_ncu_traffic.json: kernel -> {shape, dram_bytes_read, dram_bytes_write, report}); None when no capture of this
    launch shape is committed."""
    path = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    if not os.path.exists(path):
        return None
    rec = json.load(open(path)).get(name)
    if not rec or list(rec.get("shape", [])) != list(shape):
        return None
    return float(rec["dram_bytes_read"] + rec["dram_bytes_write"])


def dominant_kernel_roofline(base, B, torch):
    """The dominant kernel of the step: PLE level 0 - every expert's layer 0 -> layer 1 chained in one tcgen05 kernel plus the gate
    logits (cdcmdr_ple_chain_fwd, training form: layer-0 activation stored for the backward, dropout 0.2) = 31 % of the step's
    algorithmic FLOPs in one launch.  Timed alone with CUDA events on the launching stream; algorithmic FLOPs
    2*B*(K0*(nE*d0 + n_g) + nE*d0*d1).  When the chain kernel is not in use (CDCMDR_PLE_CHAIN=0) the concatenated-N layer-0 GEMM
    is timed instead, as in round 1.  `l0_gemm` carries that GEMM's own number either way."""
    rt = base._rt
    ws = rt.ws(B)
    lv = base._levels[0]
    D, N = base.embed_output_dim, lv.experts.G * lv.experts.dims[0]
    X = base._x_mat(ws, B)                                        # the step's own gathered-embedding buffer (same pitch)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(peaks_path))["bf16_tflops"] if os.path.exists(peaks_path) else 1590.0
    sec_l0 = _time_launch(lambda: lv.experts.fwd_layer0_only(ws, X, B), torch)
    l0 = {"kernel": "level-0 expert GEMM fwd alone (concat-N, bias+ReLU+dropout epilogue)", "achieved": 2.0 * B * D * N / sec_l0 / 1e12,
          "us_per_launch": sec_l0 * 1e6, "frac": 2.0 * B * D * N / sec_l0 / 1e12 / peak}
    if getattr(lv, "chain", False):
        d0, d1 = lv.experts.dims
        flops = 2.0 * B * (D * (N + lv.n_gcols) + lv.experts.G * d0 * d1)
        base.train()
        sec = _time_launch(lambda: base._chain_launch(ws, 0, X, B, True, True), torch)
        sec_inf = _time_launch(lambda: base._chain_launch(ws, 0, X, B, False, False), torch)
        ach = flops / sec / 1e12
        return {"kernel": "cdcmdr_ple_chain_fwd: PLE level 0, expert layers 0->1 chained in-kernel + gate logits (training form)",
                "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "traffic": _ncu_traffic("ple_chain_fwd_kernel", [B, D, lv.experts.G, d0, d1, lv.n_gcols]),
                "us_per_launch": sec * 1e6, "algorithmic_flops": flops,
                "algorithmic_bytes": B * (D * 2 + N * 2 + lv.experts.G * d1 * 2 + lv.n_gcols * 4),
                "inference_form": {"us_per_launch": sec_inf * 1e6, "achieved": flops / sec_inf / 1e12,
                                   "note": "no layer-0 activation store, no dropout: [B, nE*d0] never reaches HBM"},
                "l0_gemm": l0,
                "note": "bound by shared-memory bandwidth (profiles/r2_chain_waits.md): N=128 UMMAs read 128 B/clk of operands, the "
                        "weight-ring fills and the epilogue's activation blocks share the pipe"}
    ach = 2.0 * B * D * N / sec_l0 / 1e12
    return {"kernel": "level-0 expert GEMM fwd (concat-N, bias+ReLU+dropout epilogue)", "bound": "tensor", "achieved": ach,
            "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": _ncu_traffic("gemm_bf16_tc_kernel_l0", [B, D, N]),
            "us_per_launch": sec_l0 * 1e6, "algorithmic_flops": 2.0 * B * D * N}


def cdc_amortised(model, opt, B, ms_step, torch):
    """the amortised CDC number next to the steady-state one.  run.py:601-604: at batch size bs the affinity matrices are
    re-probed every update_interval = 1000*1024//bs steps (15 at 65 536) with k = max(1, 2*1024//bs) steps per probe; one
    `update_matrix_cdc` = 50 treatment rows + matrix A + matrix B = ~115 probes, each {k fused steps, one batched evaluation of
    all 30 domains, restore}, + update_group.  Times ONE call (after a warm call) on per-domain batches of the bench's size."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from bench_probe import Provider
    np.random.seed(SEED)
    get = Provider(field_dims(), B, 2, 11)
    k = max(1, 2 * 1024 // B)
    interval = max(1, 1000 * 1024 // B)
    model.update_matrix_cdc(get, opt, k)                                  # warm call: workspaces of every probe batch size
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    model.update_matrix_cdc(get, opt, k)
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    model.train()
    amort = interval * B / (interval * ms_step * 1e-3 + sec)
    return {"update_matrix_cdc_s": sec, "update_interval_steps": interval, "k": k, "samples_per_s": amort,
            "note": "steady-state steps + one update_matrix_cdc per update_interval (run.py:596-645); probe rows are not counted as samples"}


def embedding_bandwidth(base, x, B, torch):
    """The gather kernel alone: algorithmic bytes F*(4 + E*4 + E*s_out) per sample, two ways:
    `achieved` - the workload's own table (64 MB < the 126 MB L2) and Zipf ids: hot rows are L2 hits, so this is NOT a DRAM number;
    `uniform_big_table` - uniform ids over a 2 GiB table (>> L2): every 64-byte row comes from HBM - the worst case SURVEY 8d asks
    for, and the figure to hold against the measured HBM peak."""
    rt = base._rt
    ws = rt.ws(B)
    table = base.embedding.embedding_dict.weight
    X = ws.mat("probe.X", B, F * E, rt.act_dtype)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(peaks_path))["hbm_gbs"] if os.path.exists(peaks_path) else 6650.0
    nbytes = B * F * (4 + E * 4 + E * (2 if rt.bf16 else 4))

    def run(xi, off, tab):
        fn = lambda: rt.ops.embed_gather(xi, off, tab, None if rt.bf16 else X, X if rt.bf16 else None, B, F, E, tab.shape[0])   # noqa: E731
        return _time_launch(fn, torch, n=20)
    sec = run(x, base.embedding.offsets_dev, table)
    gbs = nbytes / sec / 1e9
    out = {"kernel": "embed_gather_fwd", "bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
           "us_per_launch": sec * 1e6, "note": "workload table (64 MB) + Zipf ids: hot rows hit L2, so this can exceed the DRAM copy peak"}
    rows_big = (2 << 30) // (E * 4)
    big = torch.empty(rows_big, E, dtype=torch.float32, device=x.device).normal_()
    per = rows_big // F
    xu = torch.randint(0, per, (B, F), dtype=torch.int32, device=x.device)
    offu = torch.arange(F, dtype=torch.int64, device=x.device) * per
    secu = run(xu, offu, big)
    out["uniform_big_table"] = {"achieved": nbytes / secu / 1e9, "frac": nbytes / secu / 1e9 / peak, "us_per_launch": secu * 1e6,
                                "table_bytes": rows_big * E * 4, "note": "uniform ids over a 2 GiB table: every row read is a DRAM access"}
    del big
    return out


_OUT = None


def emit(line):
    """The one JSON line goes to the ORIGINAL stdout; everything else any library prints to fd 1 (NCCL's version banner,
    for one) has been redirected to stderr by main()."""
    _OUT.write(json.dumps(line) + "\n")
    _OUT.flush()


def main():
    global _OUT
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("CDCMDR_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--cpu-batch", type=int, default=8192)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--trace", default=None, help="diagnostic: write per-kernel CUPTI totals of a few steps to this file")
    ap.add_argument("--no-amortised", action="store_true", help="skip the update_matrix_cdc timing behind cdc_amortised")
    ap.add_argument("--ref-budget", type=float, default=150.0, help="--impl reference: stop timing after this many seconds")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours_arm(args)


if __name__ == "__main__":
    main()
