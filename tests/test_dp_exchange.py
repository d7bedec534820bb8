"""the fused embedding exchange (csrc/dp_exchange.cu) - indices pushed to the owners, rows gathered by the owner and
stored straight into the requesters' X, row gradients pushed to the owners - against the definition (what the round-1
pack / all-to-all / unpack sequence produced).  The "ranks" are buffers inside one process: on the CPU through the host emulator
(pins the emulator), on the B200 through libcdcmdr.so with every rank's kernels on one device (bit-exact against the emulator;
the barrier is exercised with two concurrent streams).  The N-process run is tools/dp_check.py under `gpurun --gpus 2`."""
import numpy as np
import pytest
import torch

import cdcmdr_b200 as cm
split_fields = cm.parallel.split_fields
from oracle.host_abi import HostABI, f32_to_bf16


def _setup(world, B, F, E, seed, device):
    rng = np.random.default_rng(seed)
    dims = rng.integers(3, 40, size=F)
    off = np.concatenate([[0], np.cumsum(dims)[:-1]]).astype(np.int64)
    V = int(dims.sum())
    table = rng.standard_normal((V, E)).astype(np.float32)
    ranges = split_fields(F, world)
    bounds = np.concatenate([off, [V]])
    xs = [np.stack([rng.integers(0, dims[f], size=B) for f in range(F)], axis=1).astype(np.int32) for _ in range(world)]
    if B:
        xs[0][0, 0] = dims[0] + 10_000                         # one index outside the vocabulary: row reads as zeros, flag set
    dXs = [rng.standard_normal((B, F * E + 8)).astype(np.float32) for _ in range(world)]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    return dict(world=world, B=B, F=F, E=E, V=V, off=off, table=table, ranges=ranges, bounds=bounds, xs=xs, dXs=dXs, t=t,
                nf=[f1 - f0 for f0, f1 in ranges], fbound=np.array([r[0] for r in ranges] + [ranges[-1][1]], dtype=np.int32))


def _run(lib, c, device, out_bf16, stream=0, packed=0):
    """Every rank's calls in sequence (no barrier needed inside one stream) -> (X per rank, gradient inbox per owner, oob flags)"""
    W, B, F, E, t = c["world"], c["B"], c["F"], c["E"], c["t"]
    nf_max = max(max(c["nf"]), 1)
    ldx = F * E + 8
    xdt, gdt = (torch.bfloat16, torch.bfloat16) if out_bf16 else (torch.float32, torch.float32)
    ids_in = [torch.full((W * B * nf_max + 1,), -7, dtype=torch.int32, device=device) for _ in range(W)]
    X = [torch.zeros(max(B, 1), ldx, dtype=xdt, device=device) for _ in range(W)]
    G = [torch.zeros(W * B * nf_max * E + 1, dtype=gdt, device=device) for _ in range(W)]
    ptr = lambda ts: torch.tensor([x.data_ptr() for x in ts], dtype=torch.int64, device=device)
    ids_p, x_p, g_p = ptr(ids_in), ptr(X), ptr(G)
    fb = t(c["fbound"])
    oob = torch.zeros(W, dtype=torch.int32, device=device)
    keep = []
    for r in range(W):
        x = t(c["xs"][r]); keep.append(x)
        lib.dp_push_ids(x.data_ptr(), B, F, ids_p.data_ptr(), fb.data_ptr(), r, W, stream)
    for o in range(W):
        f0, f1 = c["ranges"][o]
        r0, r1 = int(c["bounds"][f0]), int(c["bounds"][f1])
        if f1 == f0:
            continue
        shard = t(c["table"][r0:r1]); keep.append(shard)
        offl = t(c["off"][f0:f1] - r0); keep.append(offl)
        lib.dp_gather_push(ids_in[o].data_ptr(), offl.data_ptr(), shard.data_ptr(), r1 - r0, x_p.data_ptr(), 1 if out_bf16 else 0, ldx,
                           f0 * E, B, f1 - f0, E, W, packed, oob[o:].data_ptr(), stream)
    for r in range(W):
        dX = t(c["dXs"][r]); keep.append(dX)
        lib.dp_push_grads(dX.data_ptr(), ldx, B, F, E, g_p.data_ptr(), 1 if out_bf16 else 0, fb.data_ptr(), r, W, stream)
    if device != "cpu":
        torch.cuda.synchronize()
    if packed:                                                 # the requester's unpack: owner o's contiguous [B, nf*E] block -> its columns
        for r in range(W):
            flat = X[r].reshape(-1)
            out = torch.zeros_like(X[r])
            for o, (f0, f1) in enumerate(c["ranges"]):
                if f1 > f0 and B:
                    out[:B, f0 * E:f1 * E] = flat[B * f0 * E:B * f1 * E].view(B, (f1 - f0) * E)
            X[r] = out
    return X, G, oob, ids_in


def _expected(c, out_bf16):
    W, B, F, E = c["world"], c["B"], c["F"], c["E"]
    Xs, Gs = [], []
    for r in range(W):
        idx = c["xs"][r].astype(np.int64) + c["off"][None, :]
        owner_lo = np.concatenate([np.repeat(c["bounds"][f0], f1 - f0) for f0, f1 in c["ranges"]])
        owner_hi = np.concatenate([np.repeat(c["bounds"][f1], f1 - f0) for f0, f1 in c["ranges"]])
        ok = (idx >= owner_lo[None, :]) & (idx < owner_hi[None, :])          # the owner checks against ITS rows
        rows = np.where(ok[..., None], c["table"][np.clip(idx, 0, c["V"] - 1)], np.float32(0)).reshape(B, F * E)
        Xs.append(rows)
    for o in range(W):
        f0, f1 = c["ranges"][o]
        Gs.append(np.concatenate([c["dXs"][r][:, f0 * E:f1 * E].reshape(-1) for r in range(W)]) if f1 > f0 else np.zeros(0, np.float32))
    if out_bf16:
        Xs = [f32_to_bf16(x).reshape(x.shape) for x in Xs]
        Gs = [f32_to_bf16(g).reshape(g.shape) for g in Gs]
    return Xs, Gs


def _bits(t):
    return t.view(torch.int16).cpu().numpy().view(np.uint16) if t.dtype == torch.bfloat16 else t.cpu().numpy()


def _check(c, X, G, oob, ids_in, out_bf16):
    W, B, F, E = c["world"], c["B"], c["F"], c["E"]
    wantX, wantG = _expected(c, out_bf16)
    for r in range(W):
        got = _bits(X[r])[:B]
        assert np.array_equal(got[:, :F * E], wantX[r]), f"X of rank {r}"
        assert not got[:, F * E:].any()                        # the pad columns stay untouched
    for o in range(W):
        n = wantG[o].size
        assert np.array_equal(_bits(G[o])[:n], wantG[o]), f"gradient inbox of owner {o}"
        assert not _bits(G[o])[n:].any()
        nf = c["nf"][o]
        want_ids = np.concatenate([c["xs"][r][:, c["ranges"][o][0]:c["ranges"][o][1]].reshape(-1) for r in range(W)]) if nf else np.zeros(0)
        got_ids = ids_in[o].cpu().numpy()
        assert np.array_equal(got_ids[:want_ids.size], want_ids) and (got_ids[want_ids.size:] == -7).all()
    if B:
        o0 = next(o for o, (f0, f1) in enumerate(c["ranges"]) if f0 <= 0 < f1)      # the out-of-range index went to field 0's owner
        assert [int(v) for v in oob.tolist()] == [1 if o == o0 else 0 for o in range(W)]


CASES = [(1, 33, 5, 16), (2, 64, 7, 16), (3, 50, 7, 8), (4, 17, 3, 32), (2, 40, 23, 4), (2, 9, 6, 64), (8, 24, 23, 16), (2, 0, 4, 16)]


@pytest.mark.parametrize("world,B,F,E", CASES)
@pytest.mark.parametrize("out_bf16", [False, True])
@pytest.mark.parametrize("packed", [0, 1])
def test_exchange_emulator_matches_definition(world, B, F, E, out_bf16, packed):
    c = _setup(world, B, F, E, 11 * world + B, "cpu")
    _check(c, *_run(HostABI(), c, "cpu", out_bf16, packed=packed), out_bf16)


def test_barrier_emulator_counts_calls():
    emu = HostABI()
    flags = np.zeros(4, dtype=np.uint64)
    seqs = np.zeros(4, dtype=np.uint64)
    ptrs = np.array([flags.ctypes.data], dtype=np.int64)
    for k in range(3):
        emu.peer_barrier(ptrs.ctypes.data, 0, 1, 2, 4, seqs.ctypes.data, 0)
    assert seqs[2] == 3 and flags[2] == 3 and seqs[[0, 1, 3]].sum() == 0
    with pytest.raises(NotImplementedError):
        emu.peer_barrier(ptrs.ctypes.data, 0, 2, 0, 4, seqs.ctypes.data, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("world,B,F,E", CASES + [(8, 4096, 23, 16), (2, 8192, 26, 32)])
@pytest.mark.parametrize("out_bf16", [False, True])
@pytest.mark.parametrize("packed", [0, 1])
def test_exchange_kernels_match_definition_gpu(world, B, F, E, out_bf16, packed):
    lib = cm._lib.load()
    c = _setup(world, B, F, E, 11 * world + B, "cuda")
    _check(c, *_run(lib, c, "cuda", out_bf16, torch.cuda.current_stream().cuda_stream, packed=packed), out_bf16)


@pytest.mark.gpu
def test_peer_barrier_two_ranks_on_two_streams_gpu():
    """Two "ranks" on one device, each on its own stream: rank 1 is held back by a long kernel; rank 0's marker written after its
    barrier must see rank 1's payload written before its barrier.  40 rounds through the same slot (sequence numbers only grow)."""
    lib = cm._lib.load()
    dev = torch.device("cuda")
    W, NS = 2, 4
    flags = [torch.zeros(NS * W, dtype=torch.int64, device=dev) for _ in range(W)]
    seqs = [torch.zeros(NS, dtype=torch.int64, device=dev) for _ in range(W)]
    fp = torch.tensor([f.data_ptr() for f in flags], dtype=torch.int64, device=dev)
    payload = torch.zeros(1, dtype=torch.int64, device=dev)
    seen = torch.zeros(40, dtype=torch.int64, device=dev)
    big = torch.randn(4096, 4096, device=dev)
    s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for k in range(40):
        with torch.cuda.stream(s1):
            _ = big @ big                                      # rank 1 arrives late
            payload.fill_(k + 1)
            lib.peer_barrier(fp.data_ptr(), 1, W, 3, NS, seqs[1].data_ptr(), s1.cuda_stream)
            lib.peer_barrier(fp.data_ptr(), 1, W, 1, NS, seqs[1].data_ptr(), s1.cuda_stream)   # rank 0 has read the payload
        with torch.cuda.stream(s0):
            lib.peer_barrier(fp.data_ptr(), 0, W, 3, NS, seqs[0].data_ptr(), s0.cuda_stream)
            seen[k:k + 1].copy_(payload)
            lib.peer_barrier(fp.data_ptr(), 0, W, 1, NS, seqs[0].data_ptr(), s0.cuda_stream)
    torch.cuda.synchronize()
    assert seen.tolist() == list(range(1, 41))
    assert seqs[0].tolist() == [0, 40, 0, 40] and seqs[1].tolist() == [0, 40, 0, 40]
