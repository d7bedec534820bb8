"""Host-side mirror of the reference's model/layer.py for the hot path: the same module tree (so `state_dict()`
keys, shapes and - for equal torch seeds - initial values match), the same constructor / forward /
get_regularization_loss surface, with every operation executed by libcdcmdr.so.

The nn.Modules below only HOLD parameters (as views into the flat arena of runtime.py); none of them has a torch
forward.  BaseModel drives the CUDA program of its subclass:
  forward(x)                      drop-in for run.py:483 (autograd-compatible: loss.backward() runs our backward)
  get_regularization_loss(device) drop-in for run.py:489 (layer.py:96-112)
  train_step(...)                 the fused step: forward + tower selection + BCE + backward + regulariser + Adam
                                  (run.py:483-492 / 635-640) with no host synchronisation - what bench.py times.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from . import _lib
from .core import Mat
from .runtime import MlpGroup, Runtime


# --------------------------------------------------------------------------------------------- parameter holders
class FeaturesEmbedding(nn.Module):
    """model/layer.py:129-157: one table over the concatenated vocabularies; offsets = exclusive cumsum."""

    def __init__(self, field_dims, embed_dim):
        super().__init__()
        fd = np.asarray(list(field_dims), dtype=np.int64)
        self.field_num = len(fd)
        self.output_dim0 = self.field_num
        self.embed_dim = embed_dim
        self.total_rows = int(fd.sum())
        from . import parallel
        shard = parallel.current_table_shard()
        if shard is None:
            self.rows_per, self.shard = self.total_rows, None
            self.embedding_dict = nn.Embedding(self.total_rows, embed_dim)
        else:
            # Row-range sharded table (parallel.sharded_table): this process allocates ONLY its own rows
            # [rank*rows_per, (rank+1)*rows_per) of the concatenated table - on the target device directly, a 500 M x 64 table never
            # exists in one piece anywhere.  state_dict() then holds the local shard.
            rank, world, device = shard
            self.rows_per = -(-self.total_rows // world)
            self.shard = (rank, world)
            self.embedding_dict = nn.Embedding(self.rows_per, embed_dim, device=device)
            with torch.no_grad():
                valid = max(0, min(self.rows_per, self.total_rows - rank * self.rows_per))
                self.embedding_dict.weight[valid:].zero_()       # padding rows past the end of the table
        self.offsets = np.array((0, *np.cumsum(fd)[:-1]), dtype=np.int64)
        self.register_buffer("offsets_dev", torch.from_numpy(self.offsets.copy()), persistent=False)


class FeaturesLinear(nn.Module):
    """model/layer.py:115-126."""

    def __init__(self, field_dims, output_dim=1):
        super().__init__()
        self.fc = nn.Linear(field_dims, output_dim, bias=True)


class MultiLayerPerceptron(nn.Module):
    """model/layer.py:178-206 (parameter layout only: Linear at layers.{s*j}, BatchNorm1d at layers.{s*j+1},
    s = 4 with bn else 3; optional Linear(., 1) last)."""

    def __init__(self, input_dim, layer_dims, dropout, output_layer=True, bn=True):
        super().__init__()
        self.layers = nn.ModuleList()
        for d in layer_dims:
            self.layers.append(nn.Linear(input_dim, d))
            if bn:
                self.layers.append(nn.BatchNorm1d(d))
            self.layers.append(nn.ReLU())
            self.layers.append(nn.Dropout(p=dropout))
            input_dim = d
        if output_layer:
            self.layers.append(nn.Linear(input_dim, 1))
        self.n_hidden, self.bn, self.has_out = len(layer_dims), bn, output_layer

    def lin_name(self, j):
        return f"layers.{(4 if self.bn else 3) * j}"

    def bn_name(self, j):
        return f"layers.{4 * j + 1}"

    def out_name(self):
        return f"layers.{(4 if self.bn else 3) * self.n_hidden}"


def mlp_group_names(prefixes, mlp: MultiLayerPerceptron, tag):
    """Arena block names + member parameter lists for a group of identically-shaped MLPs."""
    blocks, bufs = [], []
    names = {"W": [], "b": [], "gamma": [], "beta": [], "rmean": [], "rvar": []}
    for j in range(mlp.n_hidden):
        ln = mlp.lin_name(j)
        names["W"].append(f"{tag}.W{j}"); names["b"].append(f"{tag}.b{j}")
        blocks.append((f"{tag}.W{j}", [f"{p}.{ln}.weight" for p in prefixes]))
        blocks.append((f"{tag}.b{j}", [f"{p}.{ln}.bias" for p in prefixes]))
        if mlp.bn:
            bn = mlp.bn_name(j)
            names["gamma"].append(f"{tag}.gamma{j}"); names["beta"].append(f"{tag}.beta{j}")
            names["rmean"].append(f"{tag}.rmean{j}"); names["rvar"].append(f"{tag}.rvar{j}")
            blocks.append((f"{tag}.gamma{j}", [f"{p}.{bn}.weight" for p in prefixes]))
            blocks.append((f"{tag}.beta{j}", [f"{p}.{bn}.bias" for p in prefixes]))
            bufs.append((f"{tag}.rmean{j}", [f"{p}.{bn}.running_mean" for p in prefixes]))
            bufs.append((f"{tag}.rvar{j}", [f"{p}.{bn}.running_var" for p in prefixes]))
    if mlp.has_out:
        on = mlp.out_name()
        names["Wout"], names["bout"] = f"{tag}.Wout", f"{tag}.bout"
        blocks.append((f"{tag}.Wout", [f"{p}.{on}.weight" for p in prefixes]))
        blocks.append((f"{tag}.bout", [f"{p}.{on}.bias" for p in prefixes]))
    return names, blocks, bufs


# --------------------------------------------------------------------------------------------- autograd glue
class _ForwardFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x, kw, *params):
        ctx.model = model
        out = model._engine_forward(x, train=model.training, **kw)
        ctx.token = model._fwd_token
        return out.clone()

    @staticmethod
    def backward(ctx, dpred):
        model = ctx.model
        if ctx.token != model._fwd_token:
            raise RuntimeError("cdcmdr: backward() must follow the forward() it differentiates (activations are kept "
                               "in a per-batch-size workspace that the next forward overwrites)")
        grads = model._engine_backward(dpred.contiguous())
        return (None, None, None, *grads)


class _RegFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, *params):
        ctx.model = model
        return model._reg_value().to(torch.float32).reshape(1)

    @staticmethod
    def backward(ctx, gout):
        return (None, *ctx.model._reg_grads(float(gout.reshape(-1)[0])))


# --------------------------------------------------------------------------------------------- BaseModel
class BaseModel(nn.Module):
    """model/layer.py:10-112.  Subclasses define `_layout()` (arena blocks) and `_program_fwd/_program_bwd`."""

    def __init__(self, feature_dims, embed_dim, l2_reg_embedding=1e-5, l2_reg_linear=1e-5):
        super().__init__()
        self.feature_dims = feature_dims
        self.embedding = FeaturesEmbedding(feature_dims, embed_dim)
        self.embed_output_dim = self.embedding.output_dim0 * embed_dim
        self.embed_dim = embed_dim
        self.field_num = self.embedding.field_num
        self.linear = FeaturesLinear(self.embed_output_dim)
        self.is_concat_linear_cn = None
        self.regularization_weight = []            # [(parameter names, l1, l2)]
        self.l2_reg_embedding = float(l2_reg_embedding)
        self.add_regularization_weight(["embedding.embedding_dict.weight"], l2=l2_reg_embedding)
        self.add_regularization_weight(self.reg_filter("linear"), l2=l2_reg_linear)
        self._rt = None
        self._fwd_token = 0
        self._last = None
        self._table_state = None                   # (m, v) of the fused embedding Adam
        self.precision = "fp32"
        self.n_out = 1
        self.embedding_update = "dense_exact"      # or "sparse_lazy" (documented deviation)
        self.use_atten = False
        self._att = None
        self._keep_acts = True                     # False only inside a no-grad forward: activations no backward will read may be skipped

    def build_atten(self, config, dropout):
        """layer.py:58-69 (config.use_atten): parameter holders of the field self-attention block; computed by atten.AttnBlock."""
        from .atten import build_atten
        build_atten(self, config, dropout)
        self.use_atten = True

    # ---------------------------------------------------------------- regularisation bookkeeping (layer.py:86-112)
    def reg_filter(self, prefix):
        """Names under sub-module `prefix` that the reference's filter keeps: 'weight' in the RELATIVE name and 'bn'
        not in it (so BatchNorm gammas inside MultiLayerPerceptron are regularised)."""
        mod = self.get_submodule(prefix)
        return [f"{prefix}.{n}" for n, _ in mod.named_parameters() if "weight" in n and "bn" not in n]

    def add_regularization_weight(self, weight_list, l1=0.0, l2=0.0):
        if l1:
            raise NotImplementedError("l1 regularisation is never used by the reference's hot-path models")
        self.regularization_weight.append((list(weight_list), l1, float(l2)))

    # ---------------------------------------------------------------- arena construction
    def _finalize(self, blocks, buffer_blocks, precision="fp32", dropout=0.0):
        """blocks: [(block name, [parameter names])] in the order the GEMMs want them contiguous."""
        self.precision = precision
        self._blocks, self._buffer_blocks = list(blocks), list(buffer_blocks)
        named = dict(self.named_parameters())
        placed = {n for _, ns in self._blocks for n in ns}
        for n in named:
            if n not in placed and n != "embedding.embedding_dict.weight":
                self._blocks.append((n, [n]))
        self._dropout = float(dropout)
        self._build_runtime()

    def _build_runtime(self):
        named = dict(self.named_parameters())
        bufs = dict(self.named_buffers())
        device = named["embedding.embedding_dict.weight"].device
        rt = Runtime(device, self.precision)
        rt.dropout = self._dropout
        for bname, members in self._blocks:
            off0 = rt.params.n
            total = sum(named[m].numel() for m in members)
            rt.params.add(bname, (total,))
            o = off0
            for m in members:
                if m != bname:
                    rt.params.off[m] = o
                rt.params.shape[m] = tuple(named[m].shape)
                o += named[m].numel()
        for bname, numel in getattr(self, "_extra_blocks", []):        # derived operands (e.g. STAR's W_d*W_s): arena space, not parameters
            rt.params.add(bname, (numel,))
        for bname, members in self._buffer_blocks:
            off0 = rt.buffers.n
            total = sum(bufs[m].numel() for m in members)
            rt.buffers.add(bname, (total,))
            o = off0
            for m in members:
                if m != bname:
                    rt.buffers.off[m] = o
                rt.buffers.shape[m] = tuple(bufs[m].shape)
                o += bufs[m].numel()
        old = self._rt
        rt.allocate()
        with torch.no_grad():
            for _, members in self._blocks:
                for m in members:
                    v = rt.view(m)
                    v.copy_(named[m].data)
                    named[m].data = v
            for _, members in self._buffer_blocks:
                for m in members:
                    v = rt.buf_view(m)
                    v.copy_(bufs[m].data)
                    self._set_buffer(m, v)
            for names, _, l2 in self.regularization_weight:
                for m in names:
                    if m == "embedding.embedding_dict.weight" or l2 <= 0:
                        continue
                    o = rt.params.off[m]
                    rt.L2[o:o + named[m].numel()] += l2
            for m in self._absent_grads():
                o = rt.params.off[m]
                rt.present[o:o + named[m].numel()] = 0
            for bname, numel in getattr(self, "_extra_blocks", []):
                o = rt.params.off[bname]
                rt.present[o:o + numel] = 0
            rt.present_b1 = rt.present.clone()
            for m in self._single_row_absent():
                o = rt.params.off[m]
                rt.present_b1[o:o + named[m].numel()] = 0
        if old is not None:
            rt.dp = old.dp
        if old is not None and old.M is not None:
            rt.M, rt.V = old.M.to(device), old.V.to(device)
            rt.step_state = old.step_state.to(device) if old.step_state is not None else None
        if self._table_state is not None:
            self._table_state = tuple(t.to(device) for t in self._table_state)
        self._rt = rt
        self._last = None
        if getattr(self, "use_atten", False):
            from .atten import AttnBlock
            self._att = AttnBlock(self, rt)
        self._on_runtime_built()

    def _set_buffer(self, name, value):
        mod_name, _, leaf = name.rpartition(".")
        mod = self.get_submodule(mod_name) if mod_name else self
        mod._buffers[leaf] = value

    def _absent_grads(self):
        """Parameters whose .grad stays None in the reference (never reached by forward)."""
        return []

    def _single_row_absent(self):
        """Parameters the reference does not reach on a ONE-row batch: every BatchNorm is skipped there (MultiLayerPerceptron
        layer.py:202-204, STAR's partitioned norm and DNN star.py:134-135 / layer.py:290-292), so its affine parameters keep
        `.grad is None` and torch.optim.Adam leaves them and their moments alone (a last batch of size 1 happens whenever the
        dataset size is 1 modulo the batch size)."""
        names = []
        for mname, mod in self.named_modules():
            if isinstance(mod, nn.modules.batchnorm._NormBase):
                names += [f"{mname}.{leaf}" for leaf, _ in mod.named_parameters(recurse=False)]
        names += [n for n in ("shared_bn_weight", "shared_bn_bias") if hasattr(self, n)]
        return names

    def _on_runtime_built(self):
        pass

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse) if recurse is not True else super()._apply(fn)
        if getattr(self, "_rt", None) is not None:
            self._build_runtime()
        return out

    # ---------------------------------------------------------------- device checks
    def _check_device(self, x):
        rt = self._rt
        if rt is None:
            raise RuntimeError("cdcmdr: model was not finalised")
        emu = getattr(rt.ops.lib, "is_host_emulator", False)
        if rt.device.type != "cuda" and not emu:
            raise RuntimeError("cdcmdr: this model runs only on a CUDA device (sm_100a); there is no CPU fallback. "
                               "Move it with model.to('cuda').")
        if x.device != rt.device and not (x.device.type == rt.device.type == "cuda" and rt.device.index is None):
            raise RuntimeError(f"cdcmdr: input on {x.device}, model on {rt.device}")
        if x.dtype != torch.int32:
            raise TypeError("cdcmdr: sparse-field indices must be int32 (run.py:198-199)")
        if x.dim() != 2 or x.shape[1] != self.field_num:
            raise ValueError(f"cdcmdr: expected indices of shape [B, {self.field_num}]")

    # ---------------------------------------------------------------- forward / backward drivers
    def _x_mat(self, ws, B) -> Mat:
        """The gathered embeddings [B, F*E].  Models whose first layer takes its bias gradient from the weight-gradient GEMM
        (`_x_ones_col`, bf16 path) get 8 extra columns per row: column F*E holds 1.0, the rest 0 - written once, the gather
        never touches them - so that dY^T [X | 1] yields dW and db in one product."""
        rt = self._rt
        D = self.field_num * self.embed_dim
        if not (rt.bf16 and getattr(self, "_x_ones_col", False)):
            return ws.mat("X", B, D, rt.act_dtype)
        ld = D + 8
        if "X" not in ws.bufs or ws.bufs["X"].numel() < B * ld:
            ws.marks.discard("X.ones")
        X = ws.mat("X", B, ld, rt.act_dtype, zero=True)
        if "X.ones" not in ws.marks:
            X.t[:B * ld].view(B, ld)[:, D] = 1.0
            ws.marks.add("X.ones")
        return X

    def _gather(self, ws, x, B, plan_ahead=False, phase="all"):
        """phase (replicas with the field-sharded table only; DataParallel.embed_forward): "exchange" = fill X for `x` without
        planning, "ids" / "rows" = the two halves of that (the prefetch of the NEXT step's batch: indices under the forward, rows
        behind the table update), "consume" = X was prefetched by the previous step."""
        rt = self._rt
        E, F = self.embed_dim, self.field_num
        table = self.embedding.embedding_dict.weight
        if rt.bf16 and (F * E) % 8:
            raise ValueError("the bf16 tensor-core path needs field_num*embed_dim to be a multiple of 8 (TMA alignment)")
        if rt.dp is not None and rt.dp.shard:
            rt.dp.prepare_ws(ws, B)                            # peer exchange: the symmetric inboxes of this batch size (made once)
        X = self._x_mat(ws, B)
        if phase != "all" and not (rt.dp is not None and rt.dp.shard):
            raise ValueError("cdcmdr: x_next / prefetched apply to data-parallel replicas with the field-sharded table")
        if rt.dp is not None and rt.dp.shard:
            if phase == "consume" and rt.dp._pref != B:
                raise RuntimeError("cdcmdr: prefetched=True but no prefetched exchange is pending for this batch size (another forward "
                                   "ran in between, or the previous step had no x_next): call train_step without prefetched once")
            rt.dp.embed_forward(ws, x, B, X, plan_ahead, phase)   # row-sharded table: indices to the owners, rows back (parallel.py)
            if phase == "consume":
                rt.dp._pref = None
            if phase == "ids":
                return X
            if rt.bf16 and self._att_needs_x32():
                # the rows travel in bf16 on this path: a block that computes in fp32 reads them widened back
                rt.ops.cast_bf16_f32(X, ws.mat("X32", B, F * E), B, F * E)
            return X
        if rt.bf16:
            # an attention block that computes in fp32 reads the embeddings in fp32: the gather writes both copies in one pass
            x32 = ws.mat("X32", B, F * E) if self._att_needs_x32() else None
            rt.ops.embed_gather(x, self.embedding.offsets_dev, table, x32, X, B, F, E, table.shape[0])
        else:
            rt.ops.embed_gather(x, self.embedding.offsets_dev, table, X, None, B, F, E, table.shape[0])
        return X

    def _x32(self, ws, X: Mat, B) -> Mat:
        """The fp32 embeddings [B, F*E] (the gathered matrix itself on the fp32 path)."""
        return ws.mat("X32", B, self.field_num * self.embed_dim) if X.is_bf16 else X

    def _att_x(self, ws, X: Mat, B) -> Mat:
        """What the attention block reads: on the tensor-core path the gathered bf16 embeddings themselves (their [B, F*E] rows are the
        [B*F, E] token matrix - which is why models with the block do not pad X with a ones column), else the fp32 embeddings."""
        if X.is_bf16 and self._att.can_bf16 and X.ld == self.field_num * self.embed_dim:
            return X
        return self._x32(ws, X, B)

    def _att_needs_x32(self) -> bool:
        return self._att is not None and not (self._rt.bf16 and self._att.can_bf16 and not getattr(self, "_x_ones_col", False))

    def _engine_forward(self, x, train, **kw):
        """Runs the CUDA forward; returns predictions (a workspace view)."""
        self._check_device(x)
        x = x.contiguous()
        B = x.shape[0]
        rt = self._rt
        ws = rt.ws(B)
        rt.refresh_operands()
        X = self._gather(ws, x, B)
        logits, lin = self._program_fwd(ws, X, B, train, **kw)
        R, T = self._head_shape(B, **kw)
        pred = ws.get("pred", (R, T))
        rt.ops.sigmoid_select_bce(logits.t, lin, R, T, 3, None, 0, None, pred, None, None, None, None, 0.0)
        self._fwd_token += 1
        self._last = dict(ws=ws, x=x, X=X, B=B, R=R, T=T, train=train, logits=logits, lin=lin, pred=pred, kw=kw)
        if train:
            self._bump_batches_tracked(B)
        return self._shape_pred(pred[:R * T].view(R, T), **kw)

    def _shape_pred(self, pred, **kw):
        return pred

    def _global_rows(self, B, R, **kw):
        """rows the mean of the loss runs over: the routed rows R on one device, the GLOBAL batch under data-parallel replicas"""
        return self._rt.dp.global_rows(B) if self._rt.dp is not None else R

    def _head_shape(self, B, **kw):
        """(rows, columns) of the logits the program produced for a B-row batch.  Called after _program_fwd: STAR with row
        routing produces one column and only the routed rows (star.py:112-114)."""
        return B, self.n_out

    def _bump_batches_tracked(self, B):
        if B == 1:
            return
        bufs = [buf for name, buf in self.named_buffers() if name.endswith("num_batches_tracked") and self._tracks(name)]
        if bufs:
            torch._foreach_add_(bufs, 1)                     # one multi-tensor launch instead of one per BatchNorm module

    def _tracks(self, name):
        return True

    def _engine_backward(self, dpred):
        """dpred: gradient w.r.t. the (B, T) predictions.  Returns gradients for (dense parameters..., table)."""
        last = self._last
        rt, ws, B, R, T = self._rt, last["ws"], last["B"], last["R"], last["T"]
        if rt.dp is not None:
            raise NotImplementedError("cdcmdr: the autograd path (loss.backward()) is single-device; data-parallel replicas "
                                      "train through model.train_step()")
        dlogits = ws.get("dlogits", (R, T))
        dlin = self._dlin_mat(ws, B)
        rt.ops.sigmoid_bwd(last["pred"], self._unshape_dpred(dpred, R, T), dlogits, dlin, R, T)
        dX = self._program_bwd(ws, last["X"], B, last["train"], Mat(dlogits, 0, T), **last["kw"])
        table = self.embedding.embedding_dict.weight
        V, E, F = table.shape[0], self.embed_dim, self.field_num
        plan = rt.ops.embed_plan(last["x"], self.embedding.offsets_dev, B, F, V, E)
        gtab = torch.empty_like(table)
        rt.ops.embed_bwd_dense(dX, plan, B, F, E, V, gtab)
        G = rt.G.clone()
        grads = []
        single_row = set(self._single_row_absent()) if B == 1 else ()
        for name, p in self._autograd_params():
            if name == "embedding.embedding_dict.weight":
                grads.append(gtab)
            elif name in self._absent_set or (B == 1 and name in single_row):
                grads.append(None)
            else:
                o = rt.params.off[name]
                grads.append(G[o:o + p.numel()].view(p.shape))
        return grads

    def _unshape_dpred(self, dpred, B, T):
        return dpred.reshape(B, T)

    def _autograd_params(self):
        return list(self.named_parameters())

    @property
    def _absent_set(self):
        return set(self._absent_grads())

    def _call(self, x, **kw):
        if torch.is_grad_enabled():
            params = [p for _, p in self._autograd_params()]
            if any(p.requires_grad for p in params):
                return _ForwardFn.apply(self, x, kw, *params)
        self._keep_acts = False                     # pure inference: nothing will differentiate this forward
        try:
            return self._engine_forward(x, train=self.training, **kw).clone()
        finally:
            self._keep_acts = True

    # ---------------------------------------------------------------- regulariser (layer.py:96-112)
    def _reg_value(self):
        rt = self._rt
        table = self.embedding.embedding_dict.weight
        out = torch.zeros(2, dtype=torch.float64, device=rt.device)
        rt.ops.reg_l2_sum(rt.W, rt.L2, 0.0, rt.W.numel(), out[0:1])
        rt.ops.reg_l2_sum(table, None, self._l2_table(), table.numel(), out[1:2])
        return out[0] + out[1]

    def _l2_table(self):
        return sum(l2 for names, _, l2 in self.regularization_weight if "embedding.embedding_dict.weight" in names)

    def _reg_grads(self, scale):
        rt = self._rt
        table = self.embedding.embedding_dict.weight
        g = torch.empty_like(rt.W)
        rt.ops.reg_l2_grad(rt.W, rt.L2, 0.0, scale, g, False, rt.W.numel())
        gt = torch.empty_like(table)
        rt.ops.reg_l2_grad(table, None, self._l2_table(), scale, gt, False, table.numel())
        out = []
        regularised = {n for names, _, l2 in self.regularization_weight if l2 > 0 for n in names}
        for name, p in self._autograd_params():
            if name == "embedding.embedding_dict.weight":
                out.append(gt)
            elif name not in regularised:
                out.append(None)                             # not part of the regulariser: contributes no gradient (layer.py:96-112)
            else:
                o = rt.params.off[name]
                out.append(g[o:o + p.numel()].view(p.shape))
        return out

    def get_regularization_loss(self, device=None):
        self._check_device(torch.empty(0, self.field_num, dtype=torch.int32, device=self._rt.device))
        params = [p for _, p in self._autograd_params()]
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _RegFn.apply(self, *params)
        return self._reg_value().to(torch.float32).reshape(1)

    # ---------------------------------------------------------------- fused training step
    SEL_MODES = {"gather": 0, "col": 1, "mean": 2}

    def train_step(self, x, y, optimizer, mode="gather", sel=None, col=0, x_next=None, prefetched=False, **kw):
        """One pass of run.py:483-492 entirely on the device, no host synchronisation:
        gather -> model -> sigmoid -> tower selection -> BCE(mean) -> backward -> regulariser -> Adam (dense arena and
        embedding table).  `optimizer` is a cdcmdr Adam (optim.py).  Returns a dict of device tensors:
        loss (= bce + reg), bce, reg, pred (B, T).

        Replicas with the field-sharded table can pipeline the embedding exchange across steps (what a prefetching loader does for
        the batch itself): `x_next` = the NEXT step's indices - their exchange (indices to the owners, rows back) is issued behind
        this step's table update, next to the dense all-reduce and optimizer, and the next call passes `prefetched=True` to skip
        its own exchange.  The rows are gathered AFTER this step's update, so the numbers are those of the unpipelined loop."""
        self._check_device(x)
        if x_next is not None:
            self._check_device(x_next)
            if x_next.shape != x.shape:
                raise ValueError("cdcmdr: x_next must have the shape of x")
        # model.eval() + a training step is legal in the reference and HAPPENS there: cdc_test_all_domain (run.py:551) leaves the
        # model in eval mode, so most probe steps of update_matrix_cdc - and the rest of that epoch - run BatchNorm on its running
        # statistics (gradients flow through the fixed affine map) with dropout off.
        train = bool(self.training)
        rt = self._rt
        x = x.contiguous()
        B = x.shape[0]
        ws = rt.ws(B)
        optimizer.attach(self)
        rt.ensure_opt_state()
        optimizer.tick(rt)                                   # t += 1, dropout seed, Adam scalars
        rt.refresh_operands()
        X = self._gather(ws, x, B, plan_ahead=True, phase="consume" if prefetched else "all")
        if x_next is not None:
            # the next batch's indices travel to their owners now (the index inbox is free once this step's plan is built): on the
            # side stream behind the plan, under the forward
            x_next = x_next.contiguous()
            side_ = rt.side_stream()
            if side_ is not None:
                side_.wait_stream(torch.cuda.current_stream(rt.device))   # (also makes the side stream part of a graph capture)
                with torch.cuda.stream(side_):
                    self._gather(ws, x_next, B, phase="ids")
            else:
                self._gather(ws, x_next, B, phase="ids")
        # The backward plan of the embedding (sort of the step's indices into segments) only needs x: it runs on a side stream
        # next to the model program (a parallel branch of the CUDA graph) and is joined before the segment sums.
        table = self.embedding.embedding_dict.weight
        V, E, F = table.shape[0], self.embed_dim, self.field_num
        sharded = rt.dp is not None and rt.dp.shard
        plan, side = None, rt.side_stream()
        if not sharded:
            if side is not None:
                main = torch.cuda.current_stream(rt.device)
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    plan = rt.ops.embed_plan(x, self.embedding.offsets_dev, B, F, V, E)
            else:
                plan = rt.ops.embed_plan(x, self.embedding.offsets_dev, B, F, V, E)
        logits, lin = self._program_fwd(ws, X, B, train, **kw)
        R, T = self._head_shape(B, **kw)                     # R != B only when the program routed rows (STAR with x_group)
        pred = ws.get("pred", (R, T))
        psel = ws.get("psel", (B,))
        sums = ws.get("loss_sums", (4,), torch.float64)      # [bce_sum, table_sumsq, reg_dense, -]; [0:2] are per-rank partial sums
        dp = rt.dp
        n_global = self._global_rows(B, R, **kw)
        dlogits = ws.get("dlogits", (R, T))
        dlin = self._dlin_mat(ws, B)
        y, sel = self._route_targets(ws, y, sel, B, **kw)
        rt.ops.sigmoid_select_bce(logits.t, lin, R, T, self.SEL_MODES[mode], sel, col, y, pred, psel, sums[0:1], dlogits,
                                  dlin, 1.0 / max(n_global, 1))
        self._fwd_token += 1
        self._last = None
        if train:
            self._bump_batches_tracked(B)
        rt.arm_input_grad(side is not None)
        dX = self._program_bwd(ws, X, B, train, Mat(dlogits, 0, T), **kw)
        dx_event = getattr(rt, "_dx_event", None)
        rt.arm_input_grad(False)
        l2t = self._l2_table()

        def dense_update():
            rt.ops.reg_l2_sum(rt.W, rt.L2, 0.0, rt.W.numel(), sums[2:3])
            rt.ops.adam_dense(rt.W, rt.G, rt.M, rt.V, rt.L2, rt.present_b1 if rt.batch_rows(B) == 1 else rt.present, rt.W.numel(),
                              rt.step_state)

        def table_update():
            if self._table_state is None:
                self._table_state = (torch.zeros_like(table), torch.zeros_like(table))
            m, v = self._table_state
            lazy = self.embedding_update == "sparse_lazy"
            if lazy:
                rt.ops.reg_l2_sum(table, None, 1.0, table.numel(), sums[1:2], scratch="reduce_table")
            rt.ops.embed_bwd_adam(dX, plan, B, F, E, V, table, m, v, l2t, rt.step_state, None if lazy else sums[1:2], lazy=lazy)

        def fork_side():
            # The embedding backward goes to the side stream.  When the program marked its input gradient as final (PLE computes the
            # level-0 input gradient before that layer's bias / weight gradients) it starts there, under the remaining tensor-bound
            # GEMMs; otherwise after the whole backward.
            main = torch.cuda.current_stream(rt.device)
            if dx_event is not None:
                side.wait_event(dx_event)
            else:
                side.wait_stream(main)

        if sharded:
            # row gradients to the owners (all-to-all), owner-side segment sum + Adam over the owner's rows: on the side stream from
            # the moment dX is final, so the exchange runs under the rest of the dense backward; the dense gradients' all-reduce and
            # the dense arena's regulariser + Adam follow on the main stream.  (Host issue order = order on the process group's
            # communicator: the row exchange first.)
            if side is not None:
                bwd_done = torch.cuda.current_stream(rt.device).record_event()   # X's last reader (level-0 weight gradient) is behind this
                fork_side()
                with torch.cuda.stream(side):
                    dp.embed_backward(ws, dX, B, l2t, sums[1:2])
                dp.all_reduce_sum(rt.G)                      # dense gradients: sum over replicas of d(global mean loss)
                dense_update()
                if x_next is not None:
                    # the NEXT step's exchange, behind this step's table update (same stream) and next to the dense all-reduce /
                    # optimizer above: its rows are the updated ones, and the next step starts with X in place
                    with torch.cuda.stream(side):
                        side.wait_event(bwd_done)
                        self._gather(ws, x_next, B, phase="rows")
            else:
                dp.all_reduce_sum(rt.G)
                dense_update()
                dp.embed_backward(ws, dX, B, l2t, sums[1:2])
                if x_next is not None:
                    self._gather(ws, x_next, B, phase="rows")
        else:
            if dp is not None:
                raise NotImplementedError("cdcmdr: data-parallel replicas need the row-sharded table (shard_embedding=True)")
            if side is not None:
                fork_side()
                with torch.cuda.stream(side):
                    table_update()
                dense_update()
            else:
                dense_update()
                table_update()
        if side is not None:
            torch.cuda.current_stream(rt.device).wait_stream(side)      # joins the dense-optimizer branch
        if dp is not None:
            dp.all_reduce_sum(sums[0:2])                     # BCE sum and table sum-of-squares are per-rank partials
        return dict(sums=sums, pred=pred[:R * T].view(R, T), psel=psel[:R], B=n_global, l2_table=l2t)

    @staticmethod
    def step_losses(out):
        """Host-side view of a train_step result (synchronises): (loss, bce, reg) as Python floats, computed like
        run.py:484-489: float32 bce + float32 reg."""
        s = out["sums"].tolist()
        bce = np.float32(s[0] / out["B"])
        reg = np.float32(s[2] + out["l2_table"] * s[1])
        return float(np.float32(bce + reg)), float(bce), float(reg)

    def _route_targets(self, ws, y, sel, B, **kw):
        if y.dtype not in (torch.int16, torch.float32):
            y = y.to(torch.float32)
        y = y.reshape(-1).contiguous()
        if sel is not None:
            sel = sel.reshape(-1).contiguous()
            if sel.dtype != torch.int64:
                raise TypeError("tower selection indices must be int64 (run.py:229-230)")
        return y, sel

    # ---------------------------------------------------------------- subclass interface
    def _dlin_mat(self, ws, B) -> Mat | None:
        raise NotImplementedError

    def _program_fwd(self, ws, X: Mat, B, train, **kw):
        raise NotImplementedError

    def _program_bwd(self, ws, X: Mat, B, train, dlogits: Mat, **kw) -> Mat:
        raise NotImplementedError

    def forward(self, x):
        return self._call(x)

    # ---------------------------------------------------------------- towers (layer.py:35-56)
    def build_tower_output(self, n_tower, tower_input_dim, tower_dims, dropout):
        towers = nn.ModuleList(MultiLayerPerceptron(tower_input_dim, tower_dims, dropout, output_layer=True)
                               for _ in range(n_tower))
        output_layers = nn.ModuleList([nn.Sigmoid() for _ in range(n_tower)])
        return towers, None, output_layers

    def _tower_group(self, rt, tower_in_dim, tower_dims) -> MlpGroup:
        names, _, _ = mlp_group_names([f"towers.{t}" for t in range(self.n_tower)], self.towers[0], "towers")
        return MlpGroup(rt, "towers", self.n_tower, tower_in_dim, tower_dims, names, bn=True, out_layer=True, in_groups=None)


def precision_of(config, default="fp32"):
    return getattr(config, "cdcmdr_precision", default) if config is not None else default
