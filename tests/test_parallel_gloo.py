"""Data-parallel replicas + row-sharded embedding table on CPU: world_size 2 over gloo, the CUDA library
replaced by the host-memory emulator of its C-ABI.  Two ranks with B/2 rows each must reproduce what ONE process
computes on the concatenated B-row batch - the reference is a single-device program, so that is the parity statement:
global-batch BatchNorm statistics, global-mean loss, summed dense gradients, owner-side embedding update."""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cdcmdr_b200 as cm
from oracle.host_abi import HostABI
from tests.test_oracle_golden import bias_before_bn


class Cfg:
    use_atten = False; use_dcn = False; ple_n_expert_specific = 2; ple_n_expert_shared = 1; mmoe_n_expert = 3
    cdcmdr_precision = "fp32"


F, E, T, ND, DOM = 7, 4, 3, 6, 3
FD = np.array([11, 7, 13, ND, 9, 5, 8], dtype=np.int64)
L2 = dict(l2_reg_embedding=1e-3, l2_reg_linear=1e-3, l2_reg_dnn=1e-3)
ADAM = dict(lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)


def _build(kind, precision="fp32", FD=FD):
    torch.manual_seed(5)
    cfg = Cfg(); cfg.cdcmdr_precision = precision
    if kind == "ple_atten":                                   # the stock config's field self-attention block on replicas
        cfg.use_atten, cfg.atten_embed_dim, cfg.att_layer_num, cfg.att_head_num, cfg.att_res = True, 8, 2, 2, True
        return cm.PLE(FD, E, T, 2, 1, ((16, 8), (8,)), (8, 8), dropout=0.0, config=cfg, **L2)
    if kind == "ple":
        return cm.PLE(FD, E, T, 2, 1, ((16, 8), (8,)), (8, 8), dropout=0.0, config=cfg, **L2)
    if kind == "mmoe":
        return cm.MMoE(FD, E, T, 3, (16, 8), (8, 8), dropout=0.0, config=cfg, **L2)
    m = cm.CDC(FD, E, T, ND, "ple", ((16, 8), (8,)), (8, 8), DOM, dropout=0.0, config=cfg, **L2)
    m.set_groups([d % T for d in range(ND)])
    return m


def _data(B, FD=FD):
    rng = np.random.default_rng(3)
    x = np.stack([rng.integers(0, d, size=B) for d in FD], axis=1).astype(np.int32)
    y = (rng.random(B) < 0.3).astype(np.int16)
    g = rng.integers(0, T, size=B).astype(np.int64)
    return x, y, g


def _steps(model, kind, x, y, g, n):
    opt = cm.Adam(model.parameters(), **ADAM)
    model.train()
    outs = []
    for _ in range(n):
        xt, yt, gt = torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(g)
        if kind == "cdc":
            out = model.train_step(xt, yt, opt, mode="split", domain_i=None)
        else:
            out = model.train_step(xt, yt, opt, mode="gather", sel=gt)
        outs.append((out["pred"].clone().numpy(), model.step_losses(out)))
    return outs


def _worker(rank, world, port, kind, B, n_steps, path, fd=None):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cm._lib.install(HostABI())
        fd = FD if fd is None else np.asarray(fd, dtype=np.int64)
        model = _build(kind, FD=fd)
        dp = cm.parallel.attach_data_parallel(model)
        x, y, g = _data(B, FD=fd)
        lo, hi = rank * B // world, (rank + 1) * B // world
        outs = _steps(model, kind, x[lo:hi], y[lo:hi], g[lo:hi], n_steps)
        # no explicit dp.gather_table(): state_dict() itself is collective under the sharded table (a pre-hook gathers the owners'
        # rows), which is what the reference's checkpoint path calls (run.py:447-459)
        sd = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
        np.savez(os.path.join(path, f"rank{rank}.npz"), **sd,
                 **{f"pred{i}": o[0] for i, o in enumerate(outs)}, **{f"loss{i}": np.array(o[1]) for i, o in enumerate(outs)})
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("kind", ["ple", "mmoe", "cdc", "ple_atten"])
def test_two_ranks_match_one_process(kind):
    _ranks_match_one_process(kind, 2, FD)


@pytest.mark.parametrize("kind,world,fields", [("mmoe", 3, [11, 7, 13, ND, 9, 5, 8]), ("cdc", 4, [11, 7, 13, ND, 9, 5, 8]),
                                               ("ple", 4, [11, 7, 13, ND]), ("ple", 4, [11, 7, 13])])
def test_more_ranks_and_uneven_field_splits(kind, world, fields):
    """3 and 4 ranks over 7 fields (owners of 2 / 2 / 3 and 1 / 2 / 2 / 2 fields: runs of equal owners are packed per strided-batch
    copy) and 4 ranks over 4 fields (one field each) and over 3 fields (rank 0 owns no rows of the table at all); the batch splits unevenly too (96 rows over 3 ranks is even, 4 ranks 24 each)"""
    _ranks_match_one_process(kind, world, np.asarray(fields, dtype=np.int64))


def _ranks_match_one_process(kind, world, fd):
    B, n_steps = 96, 3
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_worker, args=(world, _free_port(), kind, B, n_steps, tmp, [int(v) for v in fd]), nprocs=world, join=True)
        ranks = [dict(np.load(os.path.join(tmp, f"rank{r}.npz"))) for r in range(world)]
    old = cm._lib._LIB
    cm._lib.install(HostABI())
    try:
        model = _build(kind, FD=fd)
        x, y, g = _data(B, FD=fd)
        ref = _steps(model, kind, x, y, g, n_steps)
        sd = {k: v.detach().numpy() for k, v in model.state_dict().items()}
    finally:
        cm._lib.install(old)
    for i in range(n_steps):
        pred = np.concatenate([ranks[r][f"pred{i}"] for r in range(world)], axis=0)
        assert np.abs(pred - ref[i][0]).max() <= 2e-6, (i, float(np.abs(pred - ref[i][0]).max()))
        for r in range(world):
            assert np.allclose(ranks[r][f"loss{i}"], np.array(ref[i][1]), rtol=1e-5, atol=1e-7), (i, r)
    for k, v in sd.items():
        for r in range(world):
            got = ranks[r][k]
            if k.endswith("num_batches_tracked"):
                assert int(got) == int(v)
                continue
            # replicas are identical to each other bit for bit, and equal to the single process up to summation order
            assert np.array_equal(got, ranks[0][k]), k
            kk = k[len("base_model_instance."):] if k.startswith("base_model_instance.") else k
            if bias_before_bn("mmoe" if kind == "mmoe" else "ple", kk) or kk.endswith("running_mean"):   # running_mean tracks that bias
                # the true gradient of a Linear bias in front of train-mode BatchNorm is 0: Adam turns its rounding noise
                # into +-lr steps whose sign depends on the summation order (in the reference as much as here)
                assert np.abs(got - v).max() <= 2.1e-3 * n_steps, k
                continue
            assert np.abs(got - v).max() <= 2e-5 * max(1.0, float(np.abs(v).max())), (k, float(np.abs(got - v).max()))


def test_field_split_covers_every_field_once():
    for n_fields in (1, 5, 23, 26):
        for world in (1, 2, 4, 8):
            rg = cm.parallel.split_fields(n_fields, world)
            assert len(rg) == world and rg[0][0] == 0 and rg[-1][1] == n_fields
            assert all(rg[i][1] == rg[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in rg]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("n_fields,world", [(23, 8), (23, 4), (26, 8), (5, 8), (7, 2)])
def test_field_runs_pack_like_per_owner_copies(n_fields, world):
    """the exchange buffers of the sharded table are packed one RUN of equally sized owners per strided-batch copy
    (parallel.field_runs); the result must be the per-owner layout [owner][B, fields of owner * E] the all-to-all expects"""
    B, E = 9, 4
    rg = cm.parallel.split_fields(n_fields, world)
    runs = cm.parallel.field_runs(rg)
    assert sum(n * cnt for _, n, cnt in runs) == n_fields
    assert all(n > 0 and cnt > 0 for _, n, cnt in runs)
    abi = HostABI()
    X = np.arange(B * n_fields * E, dtype=np.uint32).reshape(B, n_fields * E)
    want = np.concatenate([X[:, f0 * E:f1 * E].reshape(-1) for f0, f1 in rg])
    packed = np.zeros(B * n_fields * E, dtype=np.uint32)
    for f0, n, cnt in runs:
        abi.copy2d_batched(X.ctypes.data + 4 * f0 * E, n * E, n_fields * E, packed.ctypes.data + 4 * B * f0 * E, B * n * E, n * E, cnt, B,
                           n * E, 4, 0)
    assert np.array_equal(packed, want)
    back = np.zeros_like(X)
    for f0, n, cnt in runs:
        abi.copy2d_batched(packed.ctypes.data + 4 * B * f0 * E, B * n * E, n * E, back.ctypes.data + 4 * f0 * E, n * E, n_fields * E, cnt, B,
                           n * E, 4, 0)
    assert np.array_equal(back, X)


# ------------------------------------------------------------------------------------------------ row-range sharded table (C5)
def _rows_worker(rank, world, port, kind, B, n_steps, path, full_sd):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cm._lib.install(HostABI())
        with cm.parallel.sharded_table(rank, world, "cpu"):
            model = _build(kind)
        emb = model.base_model_instance.embedding if kind == "cdc" else model.embedding
        rows_per = emb.rows_per
        assert emb.embedding_dict.weight.shape[0] == rows_per == -(-int(FD.sum()) // world)     # only this rank's rows exist
        sd = {k: torch.from_numpy(v) for k, v in np.load(full_sd).items()}
        key = [k for k in sd if k.endswith("embedding.embedding_dict.weight")][0]
        full = sd[key]
        local = torch.zeros(rows_per, full.shape[1])
        r0, r1 = rank * rows_per, min(full.shape[0], (rank + 1) * rows_per)
        local[:r1 - r0] = full[r0:r1]
        sd[key] = local
        model.load_state_dict(sd, strict=True)
        cm.parallel.attach_data_parallel(model, shard_embedding="rows")
        x, y, g = _data(B)
        lo, hi = rank * B // world, (rank + 1) * B // world
        outs = _steps(model, kind, x[lo:hi], y[lo:hi], g[lo:hi], n_steps)
        out = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
        out["__rows"] = np.array([r0, r1])
        np.savez(os.path.join(path, f"rank{rank}.npz"), **out,
                 **{f"pred{i}": o[0] for i, o in enumerate(outs)}, **{f"loss{i}": np.array(o[1]) for i, o in enumerate(outs)})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind,world", [("ple", 2), ("cdc", 3)])
def test_row_range_sharded_table_matches_one_process(kind, world):
    """BASELINE configs[4] mechanism on CPU: every rank constructs ONLY its own row range of the table (parallel.sharded_table),
    the forward reads rows where they live, the owners update their rows from the all-gathered indices / row gradients with the
    touched-row Adam and the incrementally maintained regulariser - against ONE process with the whole table and
    embedding_update='sparse_lazy' on the concatenated batch: predictions, losses (incl. the regulariser), every rank's rows."""
    B, n_steps = 96, 3
    old = cm._lib._LIB
    cm._lib.install(HostABI())
    try:
        ref = _build(kind)
        base = ref.base_model_instance if kind == "cdc" else ref
        base.embedding_update = "sparse_lazy"
        sd0 = {k: v.detach().numpy().copy() for k, v in ref.state_dict().items()}
        with tempfile.TemporaryDirectory() as tmp:
            full_sd = os.path.join(tmp, "sd0.npz")
            np.savez(full_sd, **sd0)
            mp.spawn(_rows_worker, args=(world, _free_port(), kind, B, n_steps, tmp, full_sd), nprocs=world, join=True)
            ranks = [dict(np.load(os.path.join(tmp, f"rank{r}.npz"))) for r in range(world)]
        x, y, g = _data(B)
        want = _steps(ref, kind, x, y, g, n_steps)
        sd = {k: v.detach().numpy() for k, v in ref.state_dict().items()}
    finally:
        cm._lib.install(old)
    for i in range(n_steps):
        pred = np.concatenate([ranks[r][f"pred{i}"] for r in range(world)], axis=0)
        assert np.abs(pred - want[i][0]).max() <= 2e-6, (i, float(np.abs(pred - want[i][0]).max()))
        for r in range(world):
            assert np.allclose(ranks[r][f"loss{i}"], np.array(want[i][1]), rtol=1e-5, atol=1e-7), (i, r, ranks[r][f"loss{i}"], want[i][1])
    key = [k for k in sd if k.endswith("embedding.embedding_dict.weight")][0]
    for r in range(world):
        r0, r1 = ranks[r]["__rows"]
        assert np.abs(ranks[r][key][:r1 - r0] - sd[key][r0:r1]).max() <= 2e-6, r


# ------------------------------------------------------------------------------------------------ row-routed STAR on replicas
def _star_build():
    torch.manual_seed(5)
    cfg = Cfg(); cfg.cdcmdr_precision = "fp32"
    m = cm.STAR(FD, E, T, (16, 8), domain_idx=DOM, dropout=0.0, config=cfg, **L2)
    with torch.no_grad():
        m.shared_bn_weight.uniform_(0.5, 1.5); m.shared_bn_bias.normal_(0, 0.1)
    return m


def _star_groups(B):
    rng = np.random.default_rng(9)
    g = rng.integers(-1, T, size=B).astype(np.int64)          # -1: rows outside every tower are dropped (star.py:85-87)
    g[: B // 2][g[: B // 2] == 2] = 0                         # tower 2 has rows on the second rank only
    return g


def _star_steps(model, x, y, g, n):
    opt = cm.Adam(model.parameters(), **ADAM)
    model.train()
    outs = []
    for _ in range(n):
        out = model.train_step(torch.from_numpy(x), torch.from_numpy(y), opt, mode="col", col=0, x_group=torch.from_numpy(g).view(-1, 1))
        outs.append((out["pred"].clone().numpy().reshape(-1), model.step_losses(out)))
    return outs


def _star_worker(rank, world, port, B, n_steps, path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cm._lib.install(HostABI())
        model = _star_build()
        cm.parallel.attach_data_parallel(model)
        x, y, _ = _data(B)
        g = _star_groups(B)
        lo, hi = rank * B // world, (rank + 1) * B // world
        outs = _star_steps(model, x[lo:hi], y[lo:hi], g[lo:hi], n_steps)
        sd = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
        np.savez(os.path.join(path, f"rank{rank}.npz"), **sd,
                 **{f"pred{i}": o[0] for i, o in enumerate(outs)}, **{f"loss{i}": np.array(o[1]) for i, o in enumerate(outs)})
    finally:
        dist.destroy_process_group()


def test_row_routed_star_on_two_replicas_matches_one_process():
    """STAR's row-routed mode (star.py:84-114: tower t sees only the rows of group t) on data-parallel replicas: every rank
    partitions its own rows, a tower's partitioned-norm / BatchNorm statistics and the loss mean run over the tower's rows on ALL
    ranks (one tower has rows on one rank only, some rows belong to no tower) - against one process on the concatenated batch."""
    B, n_steps, world = 96, 3, 2
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_star_worker, args=(world, _free_port(), B, n_steps, tmp), nprocs=world, join=True)
        ranks = [dict(np.load(os.path.join(tmp, f"rank{r}.npz"))) for r in range(world)]
    old = cm._lib._LIB
    cm._lib.install(HostABI())
    try:
        model = _star_build()
        x, y, _ = _data(B)
        g = _star_groups(B)
        ref = _star_steps(model, x, y, g, n_steps)
        sd = {k: v.detach().numpy() for k, v in model.state_dict().items()}
    finally:
        cm._lib.install(old)

    def order(gv):                                            # sample indices in partition order
        return np.concatenate([np.flatnonzero(gv == t) for t in range(T)])
    whole = order(g)
    for i in range(n_steps):
        want = dict(zip(whole.tolist(), ref[i][0].tolist()))
        for r in range(world):
            lo, hi = r * B // world, (r + 1) * B // world
            mine = lo + order(g[lo:hi])
            got = ranks[r][f"pred{i}"]
            assert len(got) == len(mine)
            assert max(abs(got[k] - want[int(s)]) for k, s in enumerate(mine)) <= 2e-6
            assert np.allclose(ranks[r][f"loss{i}"], np.array(ref[i][1]), rtol=1e-5, atol=1e-7), (i, r)
    for k, v in sd.items():
        for r in range(world):
            got = ranks[r][k]
            if k.endswith("num_batches_tracked"):
                assert int(got) == int(v), k
                continue
            assert np.array_equal(got, ranks[0][k]), k
            if bias_before_bn("star", k) or k.endswith("running_mean"):
                assert np.abs(got - v).max() <= 2.1e-3 * n_steps, k
                continue
            assert np.abs(got - v).max() <= 2e-5 * max(1.0, float(np.abs(v).max())), (k, float(np.abs(got - v).max()))


# ---------------------------------------------------------------- the next step's exchange issued behind this step's table update
def _worker_prefetch(rank, world, port, kind, B, path, fd=None):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cm._lib.install(HostABI())
        fd = FD if fd is None else np.asarray(fd, dtype=np.int64)
        rng = np.random.default_rng(17)
        xs = [np.stack([rng.integers(0, d, size=B) for d in fd], axis=1).astype(np.int32) for _ in range(4)]
        ys = [(rng.random(B) < 0.3).astype(np.int16) for _ in range(4)]
        gs = [rng.integers(0, T, size=B).astype(np.int64) for _ in range(4)]
        lo, hi = rank * B // world, (rank + 1) * B // world
        res = {}
        for variant in ("plain", "pipelined"):
            model = _build(kind, FD=fd)
            cm.parallel.attach_data_parallel(model)
            opt = cm.Adam(model.parameters(), **ADAM)
            model.train()
            for k in range(4):
                xt, yt, gt = (torch.from_numpy(a[k][lo:hi]) for a in (xs, ys, gs))
                kw = dict(mode="split", domain_i=None) if kind == "cdc" else dict(mode="gather", sel=gt)
                if variant == "pipelined":                      # step k exchanges batch k+1 behind its own table update
                    kw.update(x_next=torch.from_numpy(xs[k + 1][lo:hi]) if k < 3 else None, prefetched=k > 0)
                out = model.train_step(xt, yt, opt, **kw)
                res[f"{variant}.pred{k}"] = out["pred"].clone().numpy()
                res[f"{variant}.loss{k}"] = np.array(model.step_losses(out))
            for k_, v in model.state_dict().items():
                res[f"{variant}.{k_}"] = v.detach().numpy().copy()
            if variant == "pipelined":                          # a forward in between invalidates a pending prefetch: the next consume refuses
                xt, yt, gt = (torch.from_numpy(a[0][lo:hi]) for a in (xs, ys, gs))
                kw = dict(mode="split", domain_i=None) if kind == "cdc" else dict(mode="gather", sel=gt)
                model.train_step(xt, yt, opt, x_next=xt, **kw)
                with torch.no_grad():
                    model(xt)
                try:
                    model.train_step(xt, yt, opt, prefetched=True, **kw)
                    res["refused"] = np.array(0)
                except RuntimeError:
                    res["refused"] = np.array(1)
        np.savez(os.path.join(path, f"rank{rank}.npz"), **res)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind,world,fields", [("ple", 2, None), ("cdc", 3, None), ("mmoe", 2, None), ("ple", 4, [11, 7, 13])])
def test_pipelined_exchange_is_bit_identical_to_the_plain_loop(kind, world, fields):
    """train_step(x_next=...) / train_step(prefetched=True): the next batch's index / row exchange is issued behind this step's table
    update.  Four steps over four different batches: predictions, losses and every parameter are BIT identical to the plain loop
    (the prefetched rows are the updated ones), and a forward in between makes the next `prefetched=True` raise.  The 4-rank case has
    three fields: rank 0 owns no rows at all and still takes part in every phase."""
    B = 96
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_worker_prefetch, args=(world, _free_port(), kind, B, tmp, fields), nprocs=world, join=True)
        ranks = [dict(np.load(os.path.join(tmp, f"rank{r}.npz"))) for r in range(world)]
    for r in range(world):
        keys = [k[len("plain."):] for k in ranks[r] if k.startswith("plain.")]
        assert len(keys) > 10
        for k in keys:
            assert np.array_equal(ranks[r]["plain." + k], ranks[r]["pipelined." + k], equal_nan=True), (r, k)
        assert int(ranks[r]["refused"]) == 1


# ---------------------------------------------------------------- the tensor-core path's host wiring on replicas (bf16 rows / row gradients)
E16 = 8                                                        # field_num * embed_dim must be a multiple of 8 on the bf16 path


def _build_bf16(kind):
    torch.manual_seed(5)
    cfg = Cfg(); cfg.cdcmdr_precision = "bf16"
    if kind == "ple":
        return cm.PLE(FD, E16, T, 2, 1, ((16, 8), (8,)), (8, 8), dropout=0.0, config=cfg, **L2)
    m = cm.CDC(FD, E16, T, ND, "ple", ((16, 8), (8,)), (8, 8), DOM, dropout=0.0, config=cfg, **L2)
    m.set_groups([d % T for d in range(ND)])
    return m


def _worker_bf16(rank, world, port, kind, B, n_steps, path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cm._lib.install(HostABI())
        model = _build_bf16(kind)
        cm.parallel.attach_data_parallel(model)
        x, y, g = _data(B)
        lo, hi = rank * B // world, (rank + 1) * B // world
        outs = _steps(model, kind, x[lo:hi], y[lo:hi], g[lo:hi], n_steps)
        sd = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
        np.savez(os.path.join(path, f"rank{rank}.npz"), **sd,
                 **{f"pred{i}": o[0] for i, o in enumerate(outs)}, **{f"loss{i}": np.array(o[1]) for i, o in enumerate(outs)})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind,world", [("ple", 2), ("cdc", 3)])
def test_bf16_path_on_replicas_tracks_one_process(kind, world):
    """The bf16 path under data parallelism (bf16 rows and bf16 row gradients between replicas, cross-replica BatchNorm on bf16
    activations): N ranks against ONE process running the same bf16 path on the concatenated batch.  The only extra rounding is the
    row gradients' trip through bf16, so the bar is the bf16 path's own (2e-2 on predictions, 1e-2 relative on the loss) - and the
    replicas must still be bit-identical to each other."""
    B, n_steps = 96, 3
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_worker_bf16, args=(world, _free_port(), kind, B, n_steps, tmp), nprocs=world, join=True)
        ranks = [dict(np.load(os.path.join(tmp, f"rank{r}.npz"))) for r in range(world)]
    old = cm._lib._LIB
    cm._lib.install(HostABI())
    try:
        model = _build_bf16(kind)
        x, y, g = _data(B)
        ref = _steps(model, kind, x, y, g, n_steps)
    finally:
        cm._lib.install(old)
    for i in range(n_steps):
        pred = np.concatenate([ranks[r][f"pred{i}"] for r in range(world)], axis=0)
        assert np.abs(pred - ref[i][0]).max() <= 2e-2, (i, float(np.abs(pred - ref[i][0]).max()))
        for r in range(world):
            assert np.allclose(ranks[r][f"loss{i}"], np.array(ref[i][1]), rtol=1e-2, atol=1e-4), (i, r, ranks[r][f"loss{i}"], ref[i][1])
    for k in ranks[0]:
        if k.startswith(("pred", "loss")):
            continue
        for r in range(1, world):
            assert np.array_equal(ranks[r][k], ranks[0][k]), k
