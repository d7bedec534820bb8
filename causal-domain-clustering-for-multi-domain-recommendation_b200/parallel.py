"""One process per GPU over `torch.distributed` (NCCL on B200 / NVLink, gloo for the CPU tests): the reference is a
single-device program ("Parallelism"), so everything here is defined by what makes N ranks with B rows each
produce exactly what the reference produces on the concatenated N*B-row batch:

  * dense compute is data-parallel: every rank runs the model program on its B rows; the loss is the mean over the
    GLOBAL batch (the BCE kernel is given 1/(N*B)), the dense gradient arena is all-reduced (sum) before the fused Adam,
    so every replica applies the identical update;
  * train-mode BatchNorm statistics are over the global batch: per-feature (sum, sum of squares) and the backward's
    (sum dy, sum dy*xhat) are all-reduced between the two stages of the BatchNorm kernels (cdcmdr_bn_*_stats/_apply);
  * the embedding table is ROW-SHARDED in contiguous ranges cut at field boundaries: rank r owns the rows of fields
    [f0_r, f1_r) of the single concatenated table (layer.py:140) together with their Adam moments, and only the owner
    ever reads or writes them.  Forward: all-to-all of the index columns to their owners -> owner-side gather kernel ->
    all-to-all of the gathered rows back.  Backward: all-to-all of the row gradients to the owners -> owner-side
    sorted-segment sum over all N*B samples fused with the reference-exact Adam sweep over the owner's rows only.
    Cutting at field boundaries makes every message size a compile-time function of (B, fields) - no host
    synchronisation, so the whole step still records into one CUDA graph.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

from .core import Mat


# timing probes (never set in production): comma-separated subset of {small, dense, a2a_ids, a2a_rows} switches those exchanges off
_ABLATE = set(filter(None, os.environ.get("CDCMDR_DP_ABLATE", "").split(",")))


# ------------------------------------------------------------------------------------------------ row-range sharded construction
_TABLE_SHARD = None


class sharded_table:
    """Context manager for CONSTRUCTING a model whose embedding table is row-range sharded (BASELINE configs[4]: 500 M x 64 rows
    over 8 GPUs = 16 GB of fp32 rows per GPU): inside it FeaturesEmbedding allocates only this rank's rows
    [rank*rows_per, (rank+1)*rows_per), directly on `device`.

        with cm.parallel.sharded_table(rank, world, device):
            model = cm.CDC(field_dims, ...)                     # or STAR / PLE / ...
        model = model.to(device)
        cm.parallel.attach_data_parallel(model, shard_embedding="rows")
    """

    def __init__(self, rank, world, device):
        self.shard = (int(rank), int(world), torch.device(device))

    def __enter__(self):
        global _TABLE_SHARD
        self._old, _TABLE_SHARD = _TABLE_SHARD, self.shard
        return self

    def __exit__(self, *exc):
        global _TABLE_SHARD
        _TABLE_SHARD = self._old
        return False


def current_table_shard():
    return _TABLE_SHARD


def split_fields(n_fields: int, world: int):
    """Contiguous field ranges [(f0, f1)] per rank, sizes differing by at most one."""
    cuts = [(r * n_fields) // world for r in range(world + 1)]
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def field_runs(ranges):
    """Consecutive owners with the same number of fields: [(first field, fields per owner, owners)] - one strided-batch copy
    packs / unpacks a whole run (8 ranks over 23 fields: sizes 2,3,3,3,3,3,3,3 -> two launches instead of eight)."""
    runs = []
    for (f0, f1) in ranges:
        n = f1 - f0
        if n == 0:
            continue
        if runs and runs[-1][1] == n and runs[-1][0] + runs[-1][1] * runs[-1][2] == f0:
            runs[-1] = (runs[-1][0], n, runs[-1][2] + 1)
        else:
            runs.append((f0, n, 1))
    return runs


class PeerReduce:
    """The step's small all-reduces (cross-replica BatchNorm sums, loss sums: a few KB of doubles each, all on the critical path)
    as ONE kernel over NVLink peer memory (cdcmdr_peer_allreduce_f64, csrc/peer.cu) instead of an NCCL call apiece.  torch's
    symmetric-memory allocator only provides the buffer and its peer mappings (plumbing)."""
    MAX_N = 4096

    def __init__(self, group, device, lib):
        import torch.distributed._symmetric_memory as symm
        self.lib, self.world, self.rank = lib, dist.get_world_size(group), dist.get_rank(group)
        nbytes = int(lib.peer_allreduce_bytes(self.world, self.MAX_N))
        self.buf = symm.empty(nbytes, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, group.group_name)
        self.ptrs = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64, device=device)
        self.seq = torch.zeros(1, dtype=torch.int64, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group)                                    # nobody signals into a buffer its owner has not zeroed yet

    def usable(self, t: torch.Tensor) -> bool:
        return t.dtype == torch.float64 and t.is_contiguous() and 0 < t.numel() <= self.MAX_N and t.device == self.buf.device

    def all_reduce_sum(self, t: torch.Tensor):
        stream = torch.cuda.current_stream(t.device).cuda_stream
        self.lib.peer_allreduce_f64(self.ptrs.data_ptr(), self.rank, self.world, t.data_ptr(), t.data_ptr(), t.numel(), self.MAX_N,
                                    self.seq.data_ptr(), stream)


class PeerReduceF32:
    """The dense gradient arena's all-reduce (one fp32 vector of a fixed length per step) as a two-shot exchange over NVLink peer
    memory (cdcmdr_peer_allreduce_f32, csrc/peer.cu) instead of an NCCL ring.  Created on first use for that length (a collective:
    replicas run in lockstep)."""

    def __init__(self, group, device, lib, n):
        import torch.distributed._symmetric_memory as symm
        self.lib, self.world, self.rank, self.n = lib, dist.get_world_size(group), dist.get_rank(group), int(n)
        chunk = int(lib.peer_allreduce_f32_chunk(self.world, self.n))
        box = (self.world * chunk * 4 + 255) & ~255
        self.off_out, self.off_flags = box, 2 * box
        self.buf = symm.empty(2 * box + 2 * self.world * 8, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, group.group_name)
        bases = [int(p) for p in self.handle.buffer_ptrs]
        mk = lambda off: torch.tensor([b + off for b in bases], dtype=torch.int64, device=device)
        self.inbox, self.outbox, self.flags = mk(0), mk(self.off_out), mk(self.off_flags)
        self.my_inbox, self.my_outbox = bases[self.rank], bases[self.rank] + self.off_out
        self.seqs = torch.zeros(2, dtype=torch.int64, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group)

    def all_reduce_sum(self, t: torch.Tensor):
        stream = torch.cuda.current_stream(t.device).cuda_stream
        self.lib.peer_allreduce_f32(self.inbox.data_ptr(), self.outbox.data_ptr(), self.flags.data_ptr(), self.my_inbox, self.my_outbox,
                                    self.rank, self.world, t.data_ptr(), t.data_ptr(), self.n, self.seqs.data_ptr(), stream)


class PeerExchange:
    """The embedding exchange of one batch size over NVLink peer memory (csrc/dp_exchange.cu): every rank's index inbox, gathered
    row inbox and gradient inbox live in ONE symmetric allocation; owners store rows straight into the requesters' row inbox,
    requesters store indices / row gradients straight into the owners' inboxes; `cdcmdr_peer_barrier` orders the phases.  torch's
    symmetric-memory allocator only provides the buffer and its peer mappings (plumbing).

    Slots of the barrier: 0 = everybody finished the previous use of the inboxes, 1 = indices landed, 2 = rows landed,
    3 = row gradients landed."""
    N_SLOTS = 4

    def __init__(self, dp, B, device, lib):
        import torch.distributed._symmetric_memory as symm
        self.lib, self.B = lib, int(B)
        N, E, F = dp.world, dp.E, dp.F
        self.world, self.rank = N, dp.rank
        nf_max = max(max(dp.nf), 1)
        a256 = lambda n: (int(n) + 255) & ~255
        self.x_cols = F * E + 8                                 # room for either pitch of BaseModel._x_mat
        sizes = dict(flags=a256(self.N_SLOTS * N * 8), ids=a256(N * B * nf_max * 4), x=a256(B * self.x_cols * 4),
                     grads=a256(N * B * nf_max * E * 4))
        self.off, o = {}, 0
        for k, v in sizes.items():
            self.off[k] = o
            o += v
        self.buf = symm.empty(o, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, dp.group.group_name)
        bases = [int(p) for p in self.handle.buffer_ptrs]
        dev_ptrs = lambda key: torch.tensor([b + self.off[key] for b in bases], dtype=torch.int64, device=device)
        self.flag_ptrs, self.ids_ptrs, self.x_ptrs, self.grad_ptrs = (dev_ptrs(k) for k in ("flags", "ids", "x", "grads"))
        self.seqs = torch.zeros(self.N_SLOTS, dtype=torch.int64, device=device)
        self.fbound = torch.tensor([f0 for f0, _ in dp.ranges] + [dp.ranges[-1][1]], dtype=torch.int32, device=device)
        self.sizes = sizes
        torch.cuda.synchronize(device)
        dist.barrier(dp.group)                                  # nobody stores into a buffer its owner has not zeroed yet

    def view(self, key, dtype, numel):
        esz = torch.empty(0, dtype=dtype).element_size()
        assert numel * esz <= self.sizes[key]
        return self.buf[self.off[key]:self.off[key] + numel * esz].view(dtype)

    def barrier(self, slot, stream):
        self.lib.peer_barrier(self.flag_ptrs.data_ptr(), self.rank, self.world, slot, self.N_SLOTS, self.seqs.data_ptr(), stream)


class DataParallel:
    def __init__(self, model, group=None, shard_embedding=True):
        if not dist.is_initialized():
            raise RuntimeError("cdcmdr.parallel: torch.distributed is not initialised")
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self.model = model
        self.shard = bool(shard_embedding) and self.world > 1
        emb = model.embedding
        self.F, self.E = emb.field_num, emb.embed_dim
        off = np.asarray(emb.offsets, dtype=np.int64)
        V = int(emb.total_rows)
        self.ranges = split_fields(self.F, self.world)
        self.nf = [f1 - f0 for f0, f1 in self.ranges]
        bounds = np.concatenate([off, [V]])
        self.row_range = [(int(bounds[f0]), int(bounds[f1])) for f0, f1 in self.ranges]
        self.f0, self.f1 = self.ranges[self.rank]
        self.row0, self.row1 = self.row_range[self.rank]
        self._dev_state = None
        self._moments = None
        self.rows_override = None
        self.peer = None
        self._px_used = None
        self._pref = None                                       # batch size whose X / index inbox hold a prefetched exchange
        self._ids_pending = None                                # ... whose index inbox holds the next batch, rows still to come
        self._peer_f32 = {}
        self._px = {}                                           # batch size -> PeerExchange (or None: NCCL path)
        dev = emb.embedding_dict.weight.device
        if dev.type == "cuda" and self.world > 1 and os.environ.get("CDCMDR_PEER", "1") != "0":
            try:
                self.peer = PeerReduce(self.group, dev, model._rt.ops.lib)
            except Exception as exc:                           # no peer access between these devices: NCCL carries everything
                import warnings
                warnings.warn(f"cdcmdr: NVLink peer all-reduce unavailable ({exc!r}); small all-reduces go through NCCL")
                self.peer = None
        model._rt.dp = self
        model._dp = self
        if self.shard:
            # Only the owner updates its rows: a checkpoint (run.py:447-459 `torch.save(model.state_dict(), ...)`) must see every
            # rank's rows.  state_dict() therefore became a COLLECTIVE under the sharded table - every rank has to call it.
            self._sd_hook = model.register_state_dict_pre_hook(lambda module, prefix, keep_vars: self.gather_table())

    def _range_runs(self):
        if getattr(self, "_runs", None) is None:
            self._runs = field_runs(self.ranges)
        return self._runs

    # ---------------------------------------------------------------- collectives (plumbing only)
    def global_rows(self, B):
        """rows of the single-device batch this rank's B rows are a slice of.  A model whose towers see different row counts per
        rank (STAR with row routing) announces the tower's GLOBAL count through `rows_override` around that tower's launches."""
        if self.rows_override is not None:
            return int(self.rows_override)
        return B * self.world

    def all_reduce_sum(self, t: torch.Tensor):
        if _ABLATE and (("small" in _ABLATE and t.numel() <= 4096) or ("dense" in _ABLATE and t.numel() > 4096)):
            return                                             # timing probe only (tools/dp_ablate.sh): results are wrong
        if self.peer is not None and self.peer.usable(t):
            self.peer.all_reduce_sum(t)
            return
        if (self.peer is not None and t.dtype == torch.float32 and t.is_contiguous() and t.numel() >= (1 << 16)
                and os.environ.get("CDCMDR_PEER_DENSE", "1") != "0"):
            # the dense gradient arena: one length per model, at most two symmetric buffers are ever made
            pr = self._peer_f32.get(t.numel())
            if pr is None and len(self._peer_f32) < 2:
                pr = self._peer_f32[t.numel()] = PeerReduceF32(self.group, t.device, self.model._rt.ops.lib, t.numel())
            if pr is not None:
                pr.all_reduce_sum(t)
                return
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def _all_to_all(self, out: torch.Tensor, inp: torch.Tensor, out_splits, in_splits):
        if _ABLATE and (("a2a_ids" in _ABLATE and out.dtype == torch.int32) or ("a2a_rows" in _ABLATE and out.dtype != torch.int32)):
            return                                             # timing probe only
        if out.device.type == "cuda":
            dist.all_to_all_single(out, inp, out_splits, in_splits, group=self.group)
            return
        # gloo has no all_to_all on CPU tensors: the same exchange as pairwise sends
        outs = list(out.split(out_splits))
        ins = list(inp.split(in_splits))
        outs[self.rank].copy_(ins[self.rank])
        reqs = []
        for r in range(self.world):
            if r != self.rank and ins[r].numel():            # empty blocks (an owner without fields) are skipped on both sides
                reqs.append(dist.isend(ins[r].contiguous(), dist.get_global_rank(self.group, r), group=self.group))
        for r in range(self.world):
            if r != self.rank and outs[r].numel():
                tmp = torch.empty_like(outs[r])
                dist.recv(tmp, dist.get_global_rank(self.group, r), group=self.group)
                outs[r].copy_(tmp)
        for q in reqs:
            q.wait()

    # ---------------------------------------------------------------- sharded table
    def _state(self, device):
        if self._dev_state is None or self._dev_state["device"] != device:
            emb = self.model.embedding
            off = torch.as_tensor(np.asarray(emb.offsets[self.f0:self.f1], dtype=np.int64) - self.row0, device=device)
            if off.numel() == 0:
                off = torch.zeros(1, dtype=torch.int64, device=device)
            self._dev_state = dict(device=device, offsets_local=off)
        return self._dev_state

    def shard_view(self):
        """This rank's rows of the concatenated table (a view: the owner updates them in place)."""
        return self.model.embedding.embedding_dict.weight.data[self.row0:self.row1]

    def moments(self):
        if self._moments is None:
            sv = self.shard_view()
            self._moments = (torch.zeros_like(sv), torch.zeros_like(sv))
        return self._moments

    MAX_PEER_EXCHANGES = 3

    def peer_exchange(self, B):
        """The peer-memory exchange for batch size B, created on first use (a COLLECTIVE: replicas run in lockstep, so every rank
        meets here with the same B), or None - the NCCL all-to-all path - when peer memory is unavailable, the embedding width is
        not one the fused kernels cover, or MAX_PEER_EXCHANGES batch sizes already hold symmetric buffers."""
        if B in self._px:
            return self._px[B]
        px = None
        dev = self.model.embedding.embedding_dict.weight.device
        usable = (self.shard and dev.type == "cuda" and self.peer is not None and self.E in (4, 8, 16, 32, 64)
                  and os.environ.get("CDCMDR_PEER_EXCHANGE", "1") != "0" and not _ABLATE
                  and sum(v is not None for v in self._px.values()) < self.MAX_PEER_EXCHANGES)
        if usable:
            try:
                px = PeerExchange(self, B, dev, self.model._rt.ops.lib)
            except Exception as exc:
                import warnings
                warnings.warn(f"cdcmdr: NVLink peer exchange unavailable ({exc!r}); the embedding exchange goes through NCCL")
                px = None
        self._px[B] = px
        return px

    def prepare_ws(self, ws, B):
        """Called before the first exchange of a batch size: the symmetric buffers are made here (a collective)."""
        self.peer_exchange(B)

    def _plan_ahead(self, recv_ids, B, plan_ahead):
        """The owner-side backward plan over the received indices, on the side stream next to the model program."""
        rt = self.model._rt
        ops, N, E = rt.ops, self.world, self.E
        nf_me = self.f1 - self.f0
        st = self._state(recv_ids.device)
        self._plan, self._plan_event = None, None
        side = rt.side_stream() if plan_ahead else None
        if nf_me and plan_ahead:
            Vl = self.shard_view().shape[0]
            if side is not None:
                side.wait_stream(torch.cuda.current_stream(rt.device))
                with torch.cuda.stream(side):
                    self._plan = ops.embed_plan(recv_ids, st["offsets_local"], N * B, nf_me, Vl, E)
                    self._plan_event = side.record_event()
            else:
                self._plan = ops.embed_plan(recv_ids, st["offsets_local"], N * B, nf_me, Vl, E)

    def _embed_forward_peer(self, px, ws, x, B, X: Mat, plan_ahead, phase="all"):
        rt = self.model._rt
        ops, N, E = rt.ops, self.world, self.E
        nf_me = self.f1 - self.f0
        st = self._state(X.t.device)
        lib, stream = ops.lib, ops.stream
        recv_ids = px.view("ids", torch.int32, N * B * max(nf_me, 1))
        if phase == "consume":                                  # X and the index inbox were filled by the previous step's prefetch
            self._plan_ahead(recv_ids, B, plan_ahead)
            self._px_used = px
            return recv_ids
        if phase != "rows":
            px.barrier(0, stream)                               # every owner is done with the previous step's index inbox
            lib.dp_push_ids(x.data_ptr(), B, self.F, px.ids_ptrs.data_ptr(), px.fbound.data_ptr(), self.rank, N, stream)
            px.barrier(1, stream)
        if phase == "ids":
            return recv_ids
        self._plan, self._plan_event = None, None
        side = rt.side_stream() if (plan_ahead and phase == "all") else None
        if nf_me and plan_ahead and phase == "all":
            Vl = self.shard_view().shape[0]
            if side is not None:
                side.wait_stream(torch.cuda.current_stream(rt.device))
                with torch.cuda.stream(side):
                    self._plan = ops.embed_plan(recv_ids, st["offsets_local"], N * B, nf_me, Vl, E)
                    self._plan_event = side.record_event()
            else:
                self._plan = ops.embed_plan(recv_ids, st["offsets_local"], N * B, nf_me, Vl, E)
        # rows: every owner stores its [B, nf*E] block CONTIGUOUSLY into the requester's staging buffer (whole 512-byte warp stores
        # over NVLink; storing straight into X's rows meant runs of nf*E*2 = 96 bytes at 8 ranks and ~250 GB/s), the requester
        # unpacks the blocks into its X columns (one strided-batch copy per run of equally sized owners)
        if nf_me:
            shard = self.shard_view()
            lib.dp_gather_push(recv_ids.data_ptr(), st["offsets_local"].data_ptr(), shard.data_ptr(), shard.shape[0],
                               px.x_ptrs.data_ptr(), 1 if rt.bf16 else 0, 0, self.f0 * E, B, nf_me, E, N, 1, None, stream)
        px.barrier(2, stream)
        esz = 2 if rt.bf16 else 4
        rows_in = px.view("x", rt.act_dtype, B * self.F * E)
        for (f0, n, cnt) in self._range_runs():
            ops.copy2d_batched(rows_in.data_ptr() + esz * B * f0 * E, B * n * E, n * E, X.ptr + esz * f0 * E, n * E, X.ld, cnt, B, n * E,
                               esz)
        self._px_used = px                                      # the backward of this forward takes the same road
        return recv_ids

    def _embed_backward_peer(self, px, ws, dX: Mat, B, l2, sumsq_out):
        rt = self.model._rt
        ops, N, E = rt.ops, self.world, self.E
        nf_me = self.f1 - self.f0
        st = self._state(dX.t.device)
        lib, stream = ops.lib, ops.stream
        g16 = rt.bf16 and os.environ.get("CDCMDR_GRAD_EXCHANGE", "bf16") == "bf16"
        lib.dp_push_grads(dX.ptr, dX.ld, B, self.F, E, px.grad_ptrs.data_ptr(), 1 if g16 else 0, px.fbound.data_ptr(), self.rank, N,
                          stream)
        px.barrier(3, stream)
        if not nf_me:
            sumsq_out.zero_()
            return
        n = N * B * nf_me * E
        lazy = self.model.embedding_update == "sparse_lazy"
        if g16 and not lazy and E % 4 == 0:
            grecv = px.view("grads", torch.bfloat16, n)        # the segment sums widen the inbox on the fly: no cast pass
        elif g16:
            grecv = ws.get("dp.grad_recv", (n,), torch.float32)
            ops.cast_bf16_f32(Mat(px.view("grads", torch.bfloat16, n), 0, nf_me * E), Mat(grecv, 0, nf_me * E), N * B, nf_me * E)
        else:
            grecv = px.view("grads", torch.float32, n)
        shard = self.shard_view()
        Vl = shard.shape[0]
        plan = getattr(self, "_plan", None)
        if plan is None:
            plan = ops.embed_plan(px.view("ids", torch.int32, N * B * nf_me), st["offsets_local"], N * B, nf_me, Vl, E)
        elif self._plan_event is not None:
            torch.cuda.current_stream(rt.device).wait_event(self._plan_event)
        self._plan, self._plan_event = None, None
        m, v = self.moments()
        lazy = self.model.embedding_update == "sparse_lazy"
        if lazy:
            ops.reg_l2_sum(shard, None, 1.0, shard.numel(), sumsq_out, scratch="reduce_table")
        ops.embed_bwd_adam(Mat(grecv, 0, nf_me * E), plan, N * B, nf_me, E, Vl, shard, m, v, l2, rt.step_state,
                           None if lazy else sumsq_out, lazy=lazy)

    def embed_forward(self, ws, x, B, X: Mat, plan_ahead=False, phase="all"):
        """x: this rank's [B, F] int32 indices -> X[B, F*E] (activation dtype) through the owners of each field.
        plan_ahead (training step): the owner-side backward plan over the received indices starts on the side stream as soon as
        they have arrived and runs next to the model program.
        phase: "all" = exchange + plan; "exchange" = only fill X and the index inbox (the PREFETCH of a batch); "consume" = X and
        the inbox were prefetched: only the plan.  A training step splits the prefetch of the NEXT batch in two: "ids" (indices to
        the owners - the index inbox is free as soon as this step's plan is built, so this runs under the forward) and "rows"
        (gather + rows back + unpack into X - behind this step's table update)."""
        rt = self.model._rt
        if phase not in ("consume", "rows"):
            self._pref = None                                   # any other exchange at this batch size overwrites X and the inbox
        if phase == "rows" and self._ids_pending != B:
            raise RuntimeError("cdcmdr: the row phase of a prefetched exchange without its index phase")
        self._ids_pending = B if phase == "ids" else None
        px = self._px.get(B)
        if px is not None:
            out = self._embed_forward_peer(px, ws, x, B, X, plan_ahead, phase)
            if phase in ("exchange", "rows"):
                self._pref = B
            return out
        self._px_used = None
        ops, N, E, F = rt.ops, self.world, self.E, self.F
        nf_me = self.f1 - self.f0
        st = self._state(X.t.device)
        if phase == "consume":
            recv_ids = ws.get("dp.recv_ids", (N * B * max(nf_me, 1),), torch.int32)
            self._plan_ahead(recv_ids, B, plan_ahead)
            return recv_ids
        recv_ids = ws.get("dp.recv_ids", (N * B * max(nf_me, 1),), torch.int32)
        if phase != "rows":
            send_ids = ws.get("dp.send_ids", (B * F,), torch.int32)
            for (f0, n, cnt) in self._range_runs():           # owner o's block: [B, n] at element offset B*f0
                ops.copy2d_batched(x.data_ptr() + 4 * f0, n, F, send_ids.data_ptr() + 4 * B * f0, B * n, n, cnt, B, n, 4)
            self._all_to_all(recv_ids[:N * B * nf_me], send_ids[:B * F], [B * nf_me] * N, [B * n for n in self.nf])
        if phase == "ids":
            return recv_ids
        self._plan, self._plan_event = None, None
        side = rt.side_stream() if (plan_ahead and phase == "all") else None
        if nf_me and plan_ahead and phase == "all":
            Vl = self.shard_view().shape[0]
            if side is not None:
                side.wait_stream(torch.cuda.current_stream(rt.device))
                with torch.cuda.stream(side):
                    self._plan = ops.embed_plan(recv_ids, st["offsets_local"], N * B, nf_me, Vl, E)
                    self._plan_event = side.record_event()
            else:
                self._plan = ops.embed_plan(recv_ids, st["offsets_local"], N * B, nf_me, Vl, E)
        esz = 2 if rt.bf16 else 4
        rows_send = ws.get("dp.rows_send", (N * B * max(nf_me, 1) * E,), rt.act_dtype)
        if nf_me:
            shard = self.shard_view()
            m = Mat(rows_send, 0, nf_me * E)
            ops.embed_gather(recv_ids, st["offsets_local"], shard, None if rt.bf16 else m, m if rt.bf16 else None, N * B, nf_me, E,
                             shard.shape[0])
        rows_recv = ws.get("dp.rows_recv", (B * F * E,), rt.act_dtype)
        self._all_to_all(rows_recv[:B * F * E], rows_send[:N * B * nf_me * E], [B * n * E for n in self.nf], [B * nf_me * E] * N)
        for (f0, n, cnt) in self._range_runs():
            ops.copy2d_batched(rows_recv.data_ptr() + esz * B * f0 * E, B * n * E, n * E, X.ptr + esz * f0 * E, n * E, X.ld, cnt, B, n * E,
                               esz)
        if phase in ("exchange", "rows"):
            self._pref = B
        return recv_ids

    def embed_backward(self, ws, dX: Mat, B, l2, sumsq_out):
        """dX: fp32 [B, F*E] gradient of this rank's gathered rows -> owner-side segment sum + Adam on the owner's rows.
        sumsq_out (device double[1]): sum of squares of the owner's rows before the update (regulariser value)."""
        rt = self.model._rt
        px = self._px.get(B)
        if px is not None and getattr(self, "_px_used", None) is px:
            return self._embed_backward_peer(px, ws, dX, B, l2, sumsq_out)
        ops, N, E, F = rt.ops, self.world, self.E, self.F
        nf_me = self.f1 - self.f0
        st = self._state(dX.t.device)
        # On the tensor-core path the row gradients travel as bf16 (half the bytes of the largest exchange of the step; the owner
        # widens them back for the fp32 segment sums) - the same rounding the activations of that path carry anyway.
        g16 = rt.bf16 and os.environ.get("CDCMDR_GRAD_EXCHANGE", "bf16") == "bf16"
        gdt = torch.bfloat16 if g16 else torch.float32
        gsend = ws.get("dp.grad_send16" if g16 else "dp.grad_send", (B * F * E,), gdt)
        if g16:
            for (f0, f1) in self.ranges:                      # owner o's block [B, n*E] at element offset B*f0*E, rounded on the way
                if f1 > f0:
                    ops.cast_f32_bf16(Mat(dX.t, dX.off + f0 * E, dX.ld), Mat(gsend, B * f0 * E, (f1 - f0) * E), B, (f1 - f0) * E)
        else:
            for (f0, n, cnt) in self._range_runs():
                ops.copy2d_batched(dX.ptr + 4 * f0 * E, n * E, dX.ld, gsend.data_ptr() + 4 * B * f0 * E, B * n * E, n * E, cnt, B, n * E, 4)
        grecv = ws.get("dp.grad_recv", (N * B * max(nf_me, 1) * E,), torch.float32)
        if g16:
            grecv16 = ws.get("dp.grad_recv16", (N * B * max(nf_me, 1) * E,), torch.bfloat16)
            self._all_to_all(grecv16[:N * B * nf_me * E], gsend[:B * F * E], [B * nf_me * E] * N, [B * n * E for n in self.nf])
            if nf_me:
                ops.cast_bf16_f32(Mat(grecv16, 0, nf_me * E), Mat(grecv, 0, nf_me * E), N * B, nf_me * E)
        else:
            self._all_to_all(grecv[:N * B * nf_me * E], gsend[:B * F * E], [B * nf_me * E] * N, [B * n * E for n in self.nf])
        if not nf_me:
            sumsq_out.zero_()
            return
        shard = self.shard_view()
        Vl = shard.shape[0]
        recv_ids = ws.get("dp.recv_ids", (N * B * nf_me,), torch.int32)
        plan = getattr(self, "_plan", None)
        if plan is None:
            plan = ops.embed_plan(recv_ids, st["offsets_local"], N * B, nf_me, Vl, E)
        elif self._plan_event is not None:
            torch.cuda.current_stream(rt.device).wait_event(self._plan_event)
        self._plan, self._plan_event = None, None
        m, v = self.moments()
        lazy = self.model.embedding_update == "sparse_lazy"
        if lazy:
            ops.reg_l2_sum(shard, None, 1.0, shard.numel(), sumsq_out, scratch="reduce_table")
        ops.embed_bwd_adam(Mat(grecv, 0, nf_me * E), plan, N * B, nf_me, E, Vl, shard, m, v, l2, rt.step_state,
                           None if lazy else sumsq_out, lazy=lazy)

    def gather_table(self):
        """Make every rank's full table current (checkpointing / state_dict): broadcast each owner's rows."""
        w = self.model.embedding.embedding_dict.weight.data
        for r, (r0, r1) in enumerate(self.row_range):
            if r1 > r0:
                dist.broadcast(w[r0:r1], dist.get_global_rank(self.group, r), group=self.group)


class RowRangeParallel(DataParallel):
    """Data-parallel replicas over a ROW-RANGE sharded table in NVLink peer memory (BASELINE configs[4]).

    Rank r owns rows [r*rows_per, (r+1)*rows_per) of the single concatenated table (layer.py:140) whatever field they belong to -
    one 400 M-row field is as good as 23 equal ones - together with their Adam moments, and allocates nothing else (the model is
    built inside `sharded_table`).  Every shard lives in symmetric memory mapped into every process, so
      forward : ONE gather kernel reads each row where it lives (cdcmdr_embed_gather_peer) - no index exchange, no row exchange,
                no packing, static shapes whatever the id distribution;
      backward: the ranks all-gather their indices and bf16 row gradients (NCCL; static sizes); every owner runs the usual
                sorted-segment plan over ALL N*B samples with offsets shifted by its first row - rows outside its range fall out
                as the plan's out-of-range sentinel - and applies the touched-row ("sparse_lazy") Adam to its own rows.  The
                regulariser's table term is maintained incrementally (cdcmdr_embed_bwd_adam_sparse_lazy_reg): a full sum of
                squares per step would read the whole 16 GB shard.
    Ordering between ranks: the all-gather separates every rank's forward reads of step i from the owners' updates of step i; the
    all-reduce of the loss sums that ends every step (after the update joined the main stream) separates the updates from the
    reads of step i+1.  The dense-exact update (every row, every step: SURVEY G6) is not offered here: 96 GB of sweep per GPU per
    step at this scale."""

    def __init__(self, model, group=None):
        emb = model.embedding
        if emb.shard is None:
            raise RuntimeError("cdcmdr: build the model inside parallel.sharded_table(rank, world, device) for shard_embedding='rows'")
        super().__init__(model, group, shard_embedding=True)
        if emb.shard != (self.rank, self.world):
            raise RuntimeError(f"cdcmdr: table built as shard {emb.shard}, process is rank {self.rank} of {self.world}")
        self.shard = True                                        # also for world == 1 (the table then has one shard)
        self.rows_per, self.V = int(emb.rows_per), int(emb.total_rows)
        self.row0, self.row1 = self.rank * self.rows_per, min(self.V, (self.rank + 1) * self.rows_per)
        if hasattr(self, "_sd_hook"):
            self._sd_hook.remove()                               # state_dict() holds the local shard: nothing to gather
        model.embedding_update = "sparse_lazy"
        w = emb.embedding_dict.weight
        dev = w.device
        self._running = None
        self._peer_ptrs = None
        if dev.type == "cuda":
            import torch.distributed._symmetric_memory as symm
            buf = symm.empty((self.rows_per, self.E), dtype=torch.float32, device=dev)
            buf.copy_(w.data)
            w.data = buf                                         # the parameter now lives in symmetric memory
            self._table_handle = symm.rendezvous(buf, self.group.group_name)
            self._peer_ptrs = torch.tensor([int(p) for p in self._table_handle.buffer_ptrs], dtype=torch.int64, device=dev)
            torch.cuda.synchronize(dev)
            dist.barrier(self.group)
        off = torch.as_tensor(np.asarray(emb.offsets, dtype=np.int64) - self.row0, device=dev)
        self._offsets_local = off

    def shard_view(self):
        return self.model.embedding.embedding_dict.weight.data

    def gather_table(self):
        return None

    def prepare_ws(self, ws, B):
        pass                                                     # rows are read from their owners' shards: no exchange buffers

    def embed_forward(self, ws, x, B, X: Mat, plan_ahead=False, phase="all"):
        if phase != "all":
            raise NotImplementedError("cdcmdr: the row-range layout reads rows from their owners - there is no exchange to prefetch")
        rt = self.model._rt
        ops, N, E, F = rt.ops, self.world, self.E, self.F
        emb = self.model.embedding
        out32, out16 = (None, X) if rt.bf16 else (X, None)
        if self._peer_ptrs is not None:
            ops.lib.embed_gather_peer(x.data_ptr(), emb.offsets_dev.data_ptr(), self._peer_ptrs.data_ptr(), self.rows_per,
                                      out32.ptr if out32 is not None else None, out16.ptr if out16 is not None else None,
                                      out16.ld if out16 is not None else 0, B, F, E, self.V, None, ops.stream)
        else:
            # CPU test path (gloo, host emulator): no peer memory - assemble the table from the shards and gather locally
            full = ws.get("dp.full_table", (N * self.rows_per, E), torch.float32)
            dist.all_gather_into_tensor(full[:N * self.rows_per * E].view(N * self.rows_per, E), self.shard_view().contiguous(),
                                        group=self.group)
            ops.embed_gather(x, emb.offsets_dev, full, out32, out16, B, F, E, self.V)
        # the owners need every rank's indices for the backward: gather them now, plan on the side stream under the model program
        self._all_ids = ws.get("dp.all_ids", (N * B * F,), torch.int32)
        self._plan, self._plan_event = None, None
        if plan_ahead:
            dist.all_gather_into_tensor(self._all_ids[:N * B * F], x.reshape(-1), group=self.group)
            side = rt.side_stream()
            if side is not None:
                side.wait_stream(torch.cuda.current_stream(rt.device))
                with torch.cuda.stream(side):
                    self._plan = ops.embed_plan(self._all_ids, self._offsets_local, N * B, F, self.rows_per, E)
                    self._plan_event = side.record_event()
            else:
                self._plan = ops.embed_plan(self._all_ids, self._offsets_local, N * B, F, self.rows_per, E)
        return None

    def embed_backward(self, ws, dX: Mat, B, l2, sumsq_out):
        rt = self.model._rt
        ops, N, E, F = rt.ops, self.world, self.E, self.F
        D = F * E
        g16 = rt.bf16
        gall = ws.get("dp.grad_all", (N * B * D,), torch.float32)
        if g16:
            gsend = ws.get("dp.grad_send16", (B * D,), torch.bfloat16)
            ops.cast_f32_bf16(dX, Mat(gsend, 0, D), B, D)
            gall16 = ws.get("dp.grad_all16", (N * B * D,), torch.bfloat16)
            dist.all_gather_into_tensor(gall16[:N * B * D], gsend[:B * D], group=self.group)
            ops.cast_bf16_f32(Mat(gall16, 0, D), Mat(gall, 0, D), N * B, D)
        else:
            gsend = ws.get("dp.grad_send", (B * D,), torch.float32)
            ops.copy2d(dX.ptr, dX.ld, gsend.data_ptr(), D, B, D, 4)
            dist.all_gather_into_tensor(gall[:N * B * D], gsend[:B * D], group=self.group)
        plan = self._plan
        if plan is None:
            raise RuntimeError("cdcmdr: the row-range table trains through model.train_step (the plan is built with the forward)")
        if self._plan_event is not None:
            torch.cuda.current_stream(rt.device).wait_event(self._plan_event)
        self._plan, self._plan_event = None, None
        shard = self.shard_view()
        m, v = self.moments()
        if self._running is None:                                # sum of squares of the shard, once; maintained incrementally after
            self._running = torch.zeros(1, dtype=torch.float64, device=shard.device)
            ops.reg_l2_sum(shard, None, 1.0, shard.numel(), self._running, scratch="reduce_table")
        ops.lib.embed_bwd_adam_sparse_lazy_reg(gall.data_ptr(), D, plan.data_ptr(), E, N * B, F, E, self.rows_per, shard.data_ptr(),
                                               m.data_ptr(), v.data_ptr(), l2, rt.step_state.data_ptr(), self._running.data_ptr(),
                                               sumsq_out.data_ptr(), ops.stream)


def attach_data_parallel(model, group=None, shard_embedding=True) -> DataParallel:
    """Make `model` (a cdcmdr BaseModel, or a CDC wrapper) one replica of a data-parallel group.
    shard_embedding=True: table cut at field boundaries, all-to-all exchange (small rows, large batches: C4).
    shard_embedding="rows": row-range shards in NVLink peer memory, owner-only allocation (huge tables: C5; see RowRangeParallel)."""
    base = getattr(model, "base_model_instance", model)
    if shard_embedding == "rows":
        return RowRangeParallel(base, group)
    return DataParallel(base, group, shard_embedding)
