"""STAR (reference model/star.py:12-187) on libcdcmdr.so.

Per tower t (star.py:71-111):  h = PN_t(e) = BatchNorm(e; gamma_t*gamma_s, beta_t+beta_s)      (MDR_BatchNorm, star.py:117-181)
                               h = dropout(relu(BN_t,i(h (W_t,i * W_s,i)^T + (b_t,i + b_s,i))))  per layer (star topology FC)
                               logit_t = h (w_t * w_s)^T + (c_t + c_s) + linear(e)
  forward(x)                    every tower on every row -> (B, T)                               star.py:84, 112
  forward(x, x_group, targets)  rows are PARTITIONED by group id (stable, bit-exact order: cdcmdr_route_partition), tower t
                                runs on its own rows only -> ((n, 1) predictions in partition order, permuted targets)
                                                                                                 star.py:85-87, 107, 113-114
The effective operands W_t*W_s, b_t+b_s are derived once per step into the parameter arena (not parameters: Adam skips
them) so the towers run through the same grouped-Linear building blocks as every other MLP; the chain rule back to the
domain and shared factors is two elementwise launches per block.  `shared_dnn.bn.*` exists in the state_dict but is never
used by the reference's forward (its .grad stays None; star.py:91-96 only reads shared_dnn.linears)."""
from __future__ import annotations

import torch
from torch import nn

from .core import Mat
from .layer import BaseModel, precision_of
from .runtime import MlpGroup


class DNN(nn.Module):
    """layer.py:238-300 (parameter layout only): linears.N, bn.N (BatchNorm1d), ReLU, one Dropout."""

    def __init__(self, inputs_dim, hidden_units, activation='relu', dropout_rate=0, use_bn=True):
        super().__init__()
        if len(hidden_units) == 0:
            raise ValueError("hidden_units is empty!!")
        hu = [inputs_dim] + list(hidden_units)
        self.dropout_rate, self.use_bn = dropout_rate, use_bn
        self.dropout = nn.Dropout(dropout_rate)
        self.linears = nn.ModuleList([nn.Linear(hu[i], hu[i + 1]) for i in range(len(hu) - 1)])
        if use_bn:
            self.bn = nn.ModuleList([nn.BatchNorm1d(hu[i + 1]) for i in range(len(hu) - 1)])
        self.activation_layers = nn.ModuleList([nn.ReLU(inplace=True) for _ in range(len(hu) - 1)])


class MDR_BatchNorm(nn.BatchNorm1d):
    """star.py:117-187: the partitioned normalisation's per-domain half (weight, bias, running stats)."""


class STAR(BaseModel):
    def __init__(self, feature_dims, embed_dim, n_tower, tower_dims, domain_idx=None, dropout=0.2, config=None,
                 l2_reg_embedding=1e-5, l2_reg_linear=1e-5, l2_reg_dnn=1e-5, l2_reg_cross=1e-5, device=None):
        super().__init__(feature_dims, embed_dim, l2_reg_embedding=l2_reg_embedding, l2_reg_linear=l2_reg_linear)
        self.model_name = 'star'
        self.n_tower = self.n_out = n_tower
        self.domain_idx = domain_idx
        self.device = device
        if getattr(config, 'use_dcn', False):
            raise NotImplementedError("use_dcn=True is broken upstream (SURVEY G5) and not part of the hot path")
        if getattr(config, 'use_atten', False):
            self.build_atten(config, dropout)                  # star.py:35-36
        D, T = self.embed_output_dim, n_tower
        self.tower_dims = tuple(tower_dims)
        self.shared_bn_weight = nn.Parameter(torch.ones(D))
        self.shared_bn_bias = nn.Parameter(torch.zeros(D))
        self.domain_norm = nn.ModuleList([MDR_BatchNorm(D) for _ in range(T)])
        self.domain_dnns = nn.ModuleList([DNN(D, tower_dims, dropout_rate=dropout) for _ in range(T)])
        self.domain_dnn_linears = nn.ModuleList([nn.Linear(tower_dims[-1], 1) for _ in range(T)])
        self.shared_dnn = DNN(D, tower_dims, dropout_rate=dropout)
        self.shared_dnn_linear = nn.Linear(tower_dims[-1], 1)
        self.output_layers = nn.ModuleList([nn.Sigmoid() for _ in range(T)])
        self.add_regularization_weight(self.reg_filter("domain_dnns"), l2=l2_reg_dnn)
        self.add_regularization_weight(self.reg_filter("shared_dnn"), l2=l2_reg_dnn)

        nl = len(tower_dims)
        dims_in = [D] + list(tower_dims[:-1])
        blocks, bufs, extra = [], [], []
        names = {"W": [], "b": [], "gamma": [], "beta": [], "rmean": [], "rvar": [], "Wout": "star.Wout_eff", "bout": "star.bout_eff"}
        for i in range(nl):
            blocks.append((f"star.Wd{i}", [f"domain_dnns.{t}.linears.{i}.weight" for t in range(T)]))
            blocks.append((f"star.bd{i}", [f"domain_dnns.{t}.linears.{i}.bias" for t in range(T)]))
            blocks.append((f"star.gamma{i}", [f"domain_dnns.{t}.bn.{i}.weight" for t in range(T)]))
            blocks.append((f"star.beta{i}", [f"domain_dnns.{t}.bn.{i}.bias" for t in range(T)]))
            bufs.append((f"star.rmean{i}", [f"domain_dnns.{t}.bn.{i}.running_mean" for t in range(T)]))
            bufs.append((f"star.rvar{i}", [f"domain_dnns.{t}.bn.{i}.running_var" for t in range(T)]))
            extra += [(f"star.W_eff{i}", T * tower_dims[i] * dims_in[i]), (f"star.b_eff{i}", T * tower_dims[i])]
            names["W"].append(f"star.W_eff{i}"); names["b"].append(f"star.b_eff{i}")
            names["gamma"].append(f"star.gamma{i}"); names["beta"].append(f"star.beta{i}")
            names["rmean"].append(f"star.rmean{i}"); names["rvar"].append(f"star.rvar{i}")
        blocks.append(("star.Wdout", [f"domain_dnn_linears.{t}.weight" for t in range(T)]))
        blocks.append(("star.bdout", [f"domain_dnn_linears.{t}.bias" for t in range(T)]))
        blocks.append(("star.pn_gamma", [f"domain_norm.{t}.weight" for t in range(T)]))
        blocks.append(("star.pn_beta", [f"domain_norm.{t}.bias" for t in range(T)]))
        bufs.append(("star.pn_rmean", [f"domain_norm.{t}.running_mean" for t in range(T)]))
        bufs.append(("star.pn_rvar", [f"domain_norm.{t}.running_var" for t in range(T)]))
        extra += [("star.Wout_eff", T * tower_dims[-1]), ("star.bout_eff", T), ("star.pn_dprod", D)]
        self._names, self._extra_blocks, self._dims_in = names, extra, dims_in
        self._route = None
        self._finalize(blocks, bufs, precision=precision_of(config), dropout=dropout)

    def _absent_grads(self):
        return [n for n, _ in self.named_parameters() if n.startswith("shared_dnn.bn.")]

    def _tracks(self, name):
        return not name.startswith("shared_dnn.bn.")

    def _on_runtime_built(self):
        rt = self._rt
        self._towers = [MlpGroup(rt, f"star.t{t}", 1, self.embed_output_dim, self.tower_dims, self._names, bn=True, out_layer=True,
                                 in_groups=None, g0=t) for t in range(self.n_tower)]
        # all-rows mode (no x_group: every tower sees every row, star.py:84 / CDC over STAR): the T towers are ONE group of MLPs
        # over the T partitioned-norm outputs laid side by side [B, T*D] - one grouped GEMM / BatchNorm / head launch per layer
        # instead of T (the routed mode keeps per-tower launches: its towers see different row counts)
        self._towers_all = MlpGroup(rt, "star.all", self.n_tower, self.embed_output_dim, self.tower_dims, self._names, bn=True,
                                    out_layer=True, in_groups=None)

    # ---------------------------------------------------------------- derived operands (star.py:91-92, 100-101)
    def _derive(self):
        """W_eff[t] = W_t * W_s, b_eff[t] = b_t + b_s for every tower in one launch per block (star.py:91-92, 100-101)"""
        rt, T = self._rt, self.n_tower
        ops = rt.ops
        lo = rt.o("star.W_eff0")
        for i, d in enumerate(self.tower_dims):
            n = d * self._dims_in[i]
            ops.ewise_group(rt.w(f"star.Wd{i}"), rt.w(f"shared_dnn.linears.{i}.weight"), rt.w(f"star.W_eff{i}"), n, T, 0)
            ops.ewise_group(rt.w(f"star.bd{i}"), rt.w(f"shared_dnn.linears.{i}.bias"), rt.w(f"star.b_eff{i}"), d, T, 1)
        h = self.tower_dims[-1]
        ops.ewise_group(rt.w("star.Wdout"), rt.w("shared_dnn_linear.weight"), rt.w("star.Wout_eff"), h, T, 0)
        ops.ewise_group(rt.w("star.bdout"), rt.w("shared_dnn_linear.bias"), rt.w("star.bout_eff"), 1, T, 1)
        if rt.bf16:                                          # bf16 operand copy of the derived weights (the arena cast ran earlier)
            hi = rt.o("star.bout_eff") + T
            ops.cast_f32_bf16(Mat(rt.W, lo, hi - lo), Mat(rt.Wb, lo, hi - lo), 1, hi - lo)

    def _chain(self, active):
        """dW_eff, db_eff -> gradients of the domain and shared factors, all towers per launch (fixed tower order in the sums).
        Towers that saw no rows contribute zeros (the gradient arena was zeroed)."""
        rt, T = self._rt, self.n_tower
        ops = rt.ops
        for i, d in enumerate(self.tower_dims):
            n = d * self._dims_in[i]
            sW, sb = f"shared_dnn.linears.{i}.weight", f"shared_dnn.linears.{i}.bias"
            ge, gbe = rt.g(f"star.W_eff{i}"), rt.g(f"star.b_eff{i}")
            ops.ewise_group(ge, rt.w(sW), rt.g(f"star.Wd{i}"), n, T, 0)                       # dW_t = dW_eff_t * W_s
            ops.ewise_group(ge, rt.w(f"star.Wd{i}"), rt.g(sW), n, T, 2)                       # dW_s = sum_t dW_eff_t * W_t
            ops.ewise_group(gbe, None, rt.g(f"star.bd{i}"), d, T, 4)                          # db_t += db_eff_t (target zeroed)
            ops.ewise_group(gbe, None, rt.g(sb), d, T, 3)                                     # db_s += sum_t db_eff_t
        h = self.tower_dims[-1]
        ge, gbe = rt.g("star.Wout_eff"), rt.g("star.bout_eff")
        ops.ewise_group(ge, rt.w("shared_dnn_linear.weight"), rt.g("star.Wdout"), h, T, 0)
        ops.ewise_group(ge, rt.w("star.Wdout"), rt.g("shared_dnn_linear.weight"), h, T, 2)
        ops.ewise_group(gbe, None, rt.g("star.bdout"), 1, T, 4)
        ops.ewise_group(gbe, None, rt.g("shared_dnn_linear.bias"), 1, T, 3)

    # ---------------------------------------------------------------- routing (star.py:84-87)
    def _route_rows(self, ws, X: Mat, B, x_group):
        """Stable partition of the rows by group id.  Returns per-tower (start, count) on the HOST: the reference's boolean
        mask indexing synchronises too, and the per-tower launches need their row counts."""
        rt, T, D = self._rt, self.n_tower, self.embed_output_dim
        g = x_group.reshape(-1).contiguous()
        if g.dtype != torch.int64:
            raise TypeError("group ids must be int64 (run.py:229-230)")
        perm = ws.get("route.perm", (B,), torch.int32)
        counts = ws.get("route.counts", (T,), torch.int32)
        starts = ws.get("route.starts", (T + 1,), torch.int32)
        sc = rt.ops.scratch("route", rt.ops.lib.route_scratch_bytes(B, T))
        rt.ops.lib.route_partition(g.data_ptr(), B, T, perm.data_ptr(), counts.data_ptr(), starts.data_ptr(), sc.data_ptr(), rt.ops.stream)
        st = starts[:T + 1].tolist()                          # synchronises
        Xp = ws.mat("route.Xp", B, X.ld, X.t.dtype)
        n = st[T]
        esz = X.t.element_size()
        rt.ops.lib.permute_rows(X.ptr, X.ld, perm.data_ptr(), n, D, esz, Xp.ptr, Xp.ld, 0, rt.ops.stream)
        gcount = [st[t + 1] - st[t] for t in range(T)]
        if rt.dp is not None:
            # data-parallel replicas route their own rows; a tower's BatchNorm statistics, its one-row rule (star.py:94, 134) and
            # the mean of the loss are statements about the GLOBAL batch: the towers' row counts summed over the ranks
            import torch.distributed as dist
            cg = torch.tensor(gcount, dtype=torch.int64, device=counts.device)
            dist.all_reduce(cg, group=rt.dp.group)
            gcount = [int(v) for v in cg.tolist()]
        return dict(perm=perm, starts=st, n=n, Xp=Xp, gcount=gcount, n_global=sum(gcount))

    def _head_shape(self, B, x_group=None):
        if x_group is None:
            return B, self.n_tower
        return self._route["n"], 1

    def _permute_targets(self, targets):
        if targets is None:
            return None
        r = self._route
        t = targets.reshape(-1).contiguous()
        out = torch.empty(r["n"], dtype=t.dtype, device=t.device)
        self._rt.ops.lib.permute_rows(t.data_ptr(), 1, r["perm"].data_ptr(), r["n"], 1, t.element_size(), out.data_ptr(), 1, 0,
                                      self._rt.ops.stream)
        return out.view(-1, *targets.shape[1:]) if targets.dim() > 1 else out

    def _route_targets(self, ws, y, sel, B, x_group=None):
        y, sel = super()._route_targets(ws, y, sel, B)
        if x_group is not None:
            y = self._permute_targets(y)
        return y, sel

    def _global_rows(self, B, R, x_group=None):
        if x_group is not None and self._rt.dp is not None:
            return self._route["n_global"]
        return super()._global_rows(B, R)

    def _bump_batches_tracked(self, B):
        rows = [self._rt.batch_rows(B)] * self.n_tower if self._route is None else list(self._route["gcount"])
        for t, n in enumerate(rows):
            if n <= 1:                                        # star.py:134 (one row: PN returns its input), star.py:94; no rows: tower idle
                continue
            self.domain_norm[t].num_batches_tracked += 1
            for bn in self.domain_dnns[t].bn:
                bn.num_batches_tracked += 1

    # ---------------------------------------------------------------- program
    def _dlin_mat(self, ws, B):
        return ws.mat("dlin", B, 1)

    def _x32(self, ws, X: Mat, rows) -> Mat:
        if not X.is_bf16:
            return X
        D = self.embed_output_dim
        x32 = ws.mat("X32", rows, D)
        self._rt.ops.cast_bf16_f32(X, x32, rows, D)
        return x32

    def _pn_desc(self, ws, t, train, fwd):
        rt, D = self._rt, self.embed_output_dim
        sm = ws.get(f"star.pnsave{t}", (2, D))
        return rt.ops.bn_desc(rt.w("star.pn_gamma", t * D), rt.w("star.pn_beta", t * D),
                              rt.b("star.pn_rmean", t * D) if fwd else None, rt.b("star.pn_rvar", t * D) if fwd else None,
                              sm.data_ptr(), sm.data_ptr() + 4 * D, train, False,
                              gamma2=rt.w("shared_bn_weight"), beta2=rt.w("shared_bn_bias"))

    def forward(self, x, x_group=None, targets=None):
        if x_group is None:
            return self._call(x)
        pred = self._call(x, x_group=x_group)                 # (n, 1), rows in partition order
        if targets is None:
            return pred
        return pred, self._permute_targets(targets)           # star.py:113-114

    def _program_fwd(self, ws, X: Mat, B, train, x_group=None):
        rt, T, D = self._rt, self.n_tower, self.embed_output_dim
        ops = rt.ops
        self._derive()
        if x_group is None:
            self._route = None
            Xs, slices = X, [(0, B)] * T
            logits = ws.mat("star.logits", B, T)
        else:
            self._route = self._route_rows(ws, X, B, x_group)
            Xs = self._route["Xp"]
            st = self._route["starts"]
            slices = [(st[t], st[t + 1] - st[t]) for t in range(T)]
            logits = ws.mat("star.logits", max(self._route["n"], 1), 1)
        n_rows = B if x_group is None else self._route["n"]
        x32 = self._x32(ws, Xs, n_rows)
        lin = ws.mat("lin", max(n_rows, 1), 1)
        ops.rowdot_fwd(Xs, rt.w("linear.fc.weight"), rt.w("linear.fc.bias"), lin, n_rows, 1, D)
        if self._att is not None and n_rows > 0:
            # star.py:70-72, 103-107: atten_forward(embed_x) is per-row, `other[mask]` picks the tower's rows - the same as running
            # the block on the routed rows
            self._att.fwd(ws, x32, n_rows, lin, train)
        if x_group is None:
            hall = ws.mat("star.pn_all", B, T * D, rt.act_dtype)
            for t in range(T):
                blk = Mat(hall.t, t * D, T * D)
                if rt.batch_rows(B) == 1:                    # star.py:134-135: PN is the identity on a single row
                    ops.copy2d(Xs.ptr, Xs.ld, blk.ptr, blk.ld, B, D, Xs.t.element_size())
                else:
                    rt.bn_fwd(self._pn_desc(ws, t, train, True), x32, blk, B, D)
            return self._towers_all.fwd(ws, hall, B, train), lin
        gcount = self._route["gcount"]
        try:
            for t, (r0, n) in enumerate(slices):
                if gcount[t] == 0:                           # nobody has a row for this tower
                    continue
                if rt.dp is not None:
                    rt.dp.rows_override = gcount[t]          # the tower's rows over all replicas (a replica may hold none of them)
                xin32 = x32.rows(r0)
                if gcount[t] == 1:                           # star.py:134-135: PN is the identity on a single row
                    hin = Xs.rows(r0)
                else:
                    hin = ws.mat(f"star.pn{t}", n, D, rt.act_dtype)
                    rt.bn_fwd(self._pn_desc(ws, t, train, True), xin32, hin, n, D)
                lt = self._towers[t].fwd(ws, hin, n, train)
                ops.add2d(lt, logits.rows(r0), n, 1, False)
        finally:
            if rt.dp is not None:
                rt.dp.rows_override = None
        return logits, lin

    def _program_bwd(self, ws, X: Mat, B, train, dlogits: Mat, x_group=None):
        rt, T, D = self._rt, self.n_tower, self.embed_output_dim
        ops = rt.ops
        rt.G.zero_()                                         # idle towers and the += targets of the chain rule start from zero
        routed = x_group is not None
        if routed:
            r = self._route
            Xs, st = r["Xp"], r["starts"]
            slices = [(st[t], st[t + 1] - st[t]) for t in range(T)]
            n_rows = r["n"]
        else:
            Xs, slices, n_rows = X, [(0, B)] * T, B
        x32 = ws.mat("X32", n_rows, D) if Xs.is_bf16 else Xs
        dXs = ws.mat("star.dXs", max(n_rows, 1), D)
        first = True
        if not routed:
            hall = ws.mat("star.pn_all", B, T * D, rt.act_dtype)
            dhall = ws.mat("star.dh_all", B, T * D)
            self._towers_all.bwd(ws, hall, dlogits, B, train, dhall)
            for t in range(T):
                dh = Mat(dhall.t, t * D, T * D)
                if rt.batch_rows(B) == 1:
                    ops.add2d(dh, dXs, B, D, t > 0)
                    continue
                dprod = rt.w("star.pn_dprod")
                dz = ws.mat("star.dpn", B, D) if t > 0 else dXs
                rt.bn_bwd(self._pn_desc(ws, t, train, False), x32, None, dh, dz, dprod, rt.g("star.pn_beta", t * D), False, B, D)
                if t > 0:
                    ops.add2d(dz, dXs, B, D, True)
                ops.ewise(dprod, rt.w("shared_bn_weight"), rt.g("star.pn_gamma", t * D), D, 0)          # dgamma_t = dprod * gamma_s
                ops.ewise(dprod, rt.w("star.pn_gamma", t * D), rt.g("shared_bn_weight"), D, 2)           # dgamma_s += dprod * gamma_t
                ops.ewise(rt.g("star.pn_beta", t * D), dprod, rt.g("shared_bn_bias"), D, 3)              # dbeta_s += dbeta_t
            slices = []
        gcount = self._route["gcount"] if routed else []
        try:
            for t, (r0, n) in enumerate(slices):
                if gcount[t] == 0:
                    continue
                if rt.dp is not None:
                    rt.dp.rows_override = gcount[t]
                one = gcount[t] == 1
                dl = dlogits.rows(r0)
                hin = Xs.rows(r0) if one else ws.mat(f"star.pn{t}", n, D, rt.act_dtype)
                dh = ws.mat("star.dh", n, D)
                self._towers[t].bwd(ws, hin, dl, n, train, dh)
                tgt = dXs.rows(r0)
                if one:
                    ops.add2d(dh, tgt, n, D, False)
                else:
                    dprod = rt.w("star.pn_dprod")
                    rt.bn_bwd(self._pn_desc(ws, t, train, False), x32.rows(r0), None, dh, tgt, dprod, rt.g("star.pn_beta", t * D), False, n, D)
                    ops.ewise(dprod, rt.w("shared_bn_weight"), rt.g("star.pn_gamma", t * D), D, 0)          # dgamma_t = dprod * gamma_s
                    ops.ewise(dprod, rt.w("star.pn_gamma", t * D), rt.g("shared_bn_weight"), D, 2)           # dgamma_s += dprod * gamma_t
                    ops.ewise(rt.g("star.pn_beta", t * D), dprod, rt.g("shared_bn_bias"), D, 3)              # dbeta_s += dbeta_t
        finally:
            if rt.dp is not None:
                rt.dp.rows_override = None
        self._chain(slices)
        # wide linear (layer.py:115-126) on the rows the towers saw
        dlin = self._dlin_mat(ws, B)
        tmp = ws.mat("dX.lin", max(n_rows, 1), D)
        ops.rowdot_bwd(Xs, rt.w("linear.fc.weight"), dlin, tmp, rt.g("linear.fc.weight"), rt.g("linear.fc.bias"), n_rows, 1, D)
        ops.add2d(tmp, dXs, n_rows, D, True)
        if self._att is not None and n_rows > 0:
            self._att.bwd(ws, x32, n_rows, dlin, dXs, train)
        if not routed:
            return dXs
        dX = ws.mat("dX", B, D)
        dX.t[:B * D].zero_()                                 # rows whose group id is outside [0, T) take no gradient
        ops.lib.permute_rows(dXs.ptr, dXs.ld, self._route["perm"].data_ptr(), n_rows, D, 4, dX.ptr, dX.ld, 1, ops.stream)
        return dX
