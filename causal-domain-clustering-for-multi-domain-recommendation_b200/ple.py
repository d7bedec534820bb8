"""PLE / CGC (reference model/ple.py:9-124) on libcdcmdr.so.

Per level: all experts' first layers that read the same input run as ONE concatenated-N GEMM (level 0: every expert
and every gate reads embed_x, ple.py:54), deeper expert layers as one grouped GEMM, gate softmax + expert mixing as one
fused kernel (ple.py:106-123).  Parameter names / shapes are the reference's."""
from __future__ import annotations

import torch
from torch import nn

from .core import Mat
from .layer import BaseModel, MultiLayerPerceptron, mlp_group_names, precision_of
from .runtime import MlpGroup


class CGC(nn.Module):
    """Parameter holder with the reference's layout (ple.py:73-94)."""

    def __init__(self, cur_level, n_level, n_task, n_expert_specific, n_expert_shared, input_dims, expert_dims, dropout=0.2):
        super().__init__()
        self.cur_level, self.n_level, self.n_task = cur_level, n_level, n_task
        self.n_expert_specific, self.n_expert_shared = n_expert_specific, n_expert_shared
        self.n_expert_all = n_expert_specific * n_task + n_expert_shared
        self.experts_specific = nn.ModuleList(
            MultiLayerPerceptron(input_dims, expert_dims, dropout, output_layer=False, bn=False)
            for _ in range(n_task * n_expert_specific))
        self.experts_shared = nn.ModuleList(
            MultiLayerPerceptron(input_dims, expert_dims, dropout, output_layer=False, bn=False)
            for _ in range(n_expert_shared))
        self.gates_specific = nn.ModuleList(
            [nn.Sequential(nn.Linear(input_dims, n_expert_specific + n_expert_shared), nn.Softmax(dim=1))
             for _ in range(n_task)])
        if cur_level < n_level:
            self.gate_shared = nn.Sequential(nn.Linear(input_dims, self.n_expert_all), nn.Softmax(dim=1))


class _Level:
    pass


class PLE(BaseModel):
    def __init__(self, feature_dims, embed_dim, n_tower, n_expert_specific, n_expert_shared, expert_dims, tower_dims,
                 dropout=0.2, config=None, l2_reg_embedding=1e-5, l2_reg_linear=1e-5, l2_reg_dnn=1e-5, l2_reg_cross=1e-5,
                 model_name='ple'):
        super().__init__(feature_dims, embed_dim, l2_reg_embedding=l2_reg_embedding, l2_reg_linear=l2_reg_linear)
        self.model_name = model_name
        self.n_level = len(expert_dims)
        self.n_tower = n_tower
        self.n_out = n_tower
        if getattr(config, 'use_dcn', False):
            raise NotImplementedError("use_dcn=True is broken upstream (SURVEY G5) and not part of the hot path")
        if getattr(config, 'use_atten', False):
            self.build_atten(config, dropout)                  # ple.py:31-32 (before the CGC levels, like upstream)
        self.n_expert_specific, self.n_expert_shared = n_expert_specific, n_expert_shared
        self.expert_dims = tuple(tuple(d) for d in expert_dims)
        self.tower_dims = tuple(tower_dims)
        self.cgc_layers = nn.ModuleList(
            CGC(i + 1, self.n_level, n_tower, n_expert_specific, n_expert_shared,
                self.embed_output_dim if i == 0 else expert_dims[i - 1][-1], expert_dims[i], dropout)
            for i in range(self.n_level))
        self.towers, self.towers_linear, self.output_layers = self.build_tower_output(n_tower, expert_dims[-1][-1],
                                                                                      tower_dims, dropout)
        self.add_regularization_weight(self.reg_filter("cgc_layers"), l2=l2_reg_dnn)
        self.add_regularization_weight(self.reg_filter("towers"), l2=l2_reg_dnn)

        blocks, bufs = [], []
        self._level_names = []
        T, ns, nsh = n_tower, n_expert_specific, n_expert_shared
        for l in range(self.n_level):
            p = f"cgc_layers.{l}"
            prefixes = [f"{p}.experts_specific.{i}" for i in range(T * ns)] + [f"{p}.experts_shared.{i}" for i in range(nsh)]
            names, blk, _ = mlp_group_names(prefixes, self.cgc_layers[l].experts_specific[0], f"cgc{l}")
            gw = [f"{p}.gates_specific.{t}.0.weight" for t in range(T)]
            gb = [f"{p}.gates_specific.{t}.0.bias" for t in range(T)]
            if l + 1 < self.n_level:
                gw.append(f"{p}.gate_shared.0.weight"); gb.append(f"{p}.gate_shared.0.bias")
            if l == 0:
                gw.append("linear.fc.weight"); gb.append("linear.fc.bias")
            # expert layer-0 weights, then the gates (and the wide linear) that read the same input: one Bt matrix
            blocks += [blk[0], (f"cgc{l}.gW", gw), blk[1], (f"cgc{l}.gb", gb)] + blk[2:]
            self._level_names.append(names)
        tnames, tblk, tbufs = mlp_group_names([f"towers.{t}" for t in range(T)], self.towers[0], "towers")
        self._tower_names = tnames
        blocks += tblk
        bufs += tbufs
        # Levels >= 1: every gate reads a different input and is only 4..10 outputs wide - five launch-bound GEMMs per pass.
        # Their weights are packed once per step into ONE block-diagonal operand [n_gates*16, (T+1)*K] (gate j: rows j*16.., columns
        # of its input block) living in the arena as a derived block (not a parameter), so forward, weight-gradient and
        # input-gradient of all gates of a level are one plain GEMM each over the concatenated inputs.
        self._gpad = 16
        self._extra_blocks = []
        for l in range(1, self.n_level):
            n_gates = T if l + 1 == self.n_level else T + 1
            Kl = expert_dims[l - 1][-1]
            self._extra_blocks += [(f"cgc{l}.Wbd", n_gates * self._gpad * (T + 1) * Kl), (f"cgc{l}.bbd", n_gates * self._gpad)]
        self._finalize(blocks, bufs, precision=precision_of(config), dropout=dropout)

    # ---------------------------------------------------------------- program construction
    def _on_runtime_built(self):
        rt = self._rt
        T, ns, nsh = self.n_tower, self.n_expert_specific, self.n_expert_shared
        nE = T * ns + nsh
        self._levels = []
        for l in range(self.n_level):
            lv = _Level()
            last = l + 1 == self.n_level
            lv.K = self.embed_output_dim if l == 0 else self.expert_dims[l - 1][-1]
            lv.h = self.expert_dims[l][-1]
            lv.n_in = 1 if l == 0 else T + 1
            in_groups = [(0, 0, nE)] if l == 0 else [(t, t * ns, (t + 1) * ns) for t in range(T)] + [(T, T * ns, nE)]
            lv.experts = MlpGroup(rt, f"cgc{l}", nE, lv.K, self.expert_dims[l], self._level_names[l], bn=False,
                                  out_layer=False, in_groups=in_groups)
            lv.n_gates = T if last else T + 1
            w = ns + nsh
            n_gate_rows = T * w + (0 if last else nE)
            # gate_groups: (input block, first weight row, last weight row, first column in the logits buffer).  Level 0: every
            # gate (and the wide linear, last row) reads embed_x -> one GEMM.  Deeper levels: one GEMM per input; each gate's
            # logits start at a multiple of 8 columns so that bf16 operand views stay 16-byte aligned for TMA.
            if l == 0:
                lv.n_gcols = n_gate_rows + 1
                lv.gate_groups = [(0, 0, lv.n_gcols, 0)]
                col = [t * w for t in range(T)] + ([] if last else [T * w])
            else:
                pw = self._gpad
                if max(w, nE) > pw:
                    raise NotImplementedError("more than 16 experts per gate at a deeper CGC level")
                lv.gate_groups = [(t, t * w, (t + 1) * w, t * pw) for t in range(T)] + ([] if last else [(T, T * w, T * w + nE, T * pw)])
                lv.n_gcols = lv.n_gates * pw
                col = [j * pw for j in range(lv.n_gates)]
            lv.max_sel = w if last else nE
            n = [w] * T + ([] if last else [nE])
            sel = []
            for t in range(T):
                s = list(range(t * ns, (t + 1) * ns)) + list(range(T * ns, nE))
                sel += s + [0] * (lv.max_sel - len(s))
            if not last:
                sel += list(range(nE))
            lv.desc_t = torch.tensor(col + n + sel, dtype=torch.int32, device=rt.device)
            lv.desc = rt.ops.mix_desc(lv.n_gates, nE, lv.h, lv.max_sel, lv.desc_t, n_pairs=sum(n))
            # level 0: gates and wide linear read the same input as the experts and their weights / biases follow the experts' in
            # the arena -> their backward rides on the experts' layer-0 GEMMs as extra columns
            d0 = self.expert_dims[l][0]
            lv.fused_bwd = (l == 0 and lv.experts.can_fuse_tail() and
                            rt.o(f"cgc{l}.gW") == rt.o(self._level_names[l]["W"][0]) + nE * d0 * lv.K and
                            rt.o(f"cgc{l}.gb") == rt.o(self._level_names[l]["b"][0]) + nE * d0)
            # level 0 on the tensor-core path: expert layers 0 -> 1 and the gate logits in ONE launch (cdcmdr_ple_chain_fwd)
            lv.chain = bool(l == 0 and rt.bf16 and lv.fused_bwd and len(self.expert_dims[0]) == 2 and
                            rt.ops.ple_chain_ok(lv.K, self.expert_dims[0][0], self.expert_dims[0][1], lv.n_gcols) and
                            __import__("os").environ.get("CDCMDR_PLE_CHAIN", "1") != "0")
            if lv.fused_bwd:
                lv.experts.tail0 = lv.n_gcols
                if rt.bf16 and rt.dp is None and self._att is None:
                    # the gathered embeddings carry a column of ones: layer 0's bias gradient comes out of its weight-gradient GEMM
                    self._x_ones_col = True
                    lv.experts.x_has_ones = True
            self._levels.append(lv)
        self._towers = MlpGroup(rt, "towers", T, self.expert_dims[-1][-1], self.tower_dims, self._tower_names, bn=True,
                                out_layer=True, in_groups=None)

    # ---------------------------------------------------------------- packed gates of the deeper levels
    def _gate_blocks(self, lv):
        """(first weight row in cgc.gW, rows, packed row, input block) per gate"""
        return [(r0, r1 - r0, c0, blk) for (blk, r0, r1, c0) in lv.gate_groups]

    def _gate_runs(self, lv):
        """The gate blocks grouped into arithmetic runs: (r0, rows, c0, blk, n_blocks, d_r0, d_c0, d_blk) - every block of a run has
        `rows` rows and starts d_r0 weight rows / d_c0 packed rows / d_blk input blocks after the previous one (the T task gates
        of a level are one run; the shared gate of a non-last level is a run of its own)."""
        runs = []
        for (r0, n, c0, blk) in self._gate_blocks(lv):
            if runs:
                pr0, pn, pc0, pblk, nb, dr, dc, db = runs[-1]
                lr0, lc0, lblk = pr0 + (nb - 1) * dr, pc0 + (nb - 1) * dc, pblk + (nb - 1) * db
                step = (r0 - lr0, c0 - lc0, blk - lblk)
                if n == pn and (nb == 1 or step == (dr, dc, db)):
                    runs[-1] = (pr0, pn, pc0, pblk, nb + 1, *step)
                    continue
            runs.append((r0, n, c0, blk, 1, 0, 0, 0))
        return runs

    def _pack_gates(self, l, lv):
        rt, K = self._rt, lv.K
        ops, ldw = rt.ops, lv.n_in * lv.K
        for (r0, n, c0, blk, nb, dr, dc, db) in self._gate_runs(lv):       # runs of equally spaced, equally sized gate blocks
            ops.copy2d_batched(rt.w(f"cgc{l}.gW", r0 * K), dr * K, K, rt.w(f"cgc{l}.Wbd", c0 * ldw + blk * K), dc * ldw + db * K, ldw,
                               nb, n, K, 4)
            ops.copy2d_batched(rt.w(f"cgc{l}.gb", r0), dr, n, rt.w(f"cgc{l}.bbd", c0), dc, n, nb, 1, n, 4)
        if rt.bf16:                                              # bf16 operand copy of the packed block (the arena cast ran earlier)
            lo, n_el = rt.o(f"cgc{l}.Wbd"), lv.n_gcols * ldw
            ops.cast_f32_bf16(Mat(rt.W, lo, n_el), Mat(rt.Wb, lo, n_el), 1, n_el)

    def _gates_bwd_packed(self, ws, l, lv, xin, dLg, dxin, B):
        rt, K = self._rt, lv.K
        ops, ldw = rt.ops, lv.n_in * lv.K
        tmpb = ws.get(f"cgc{l}.dbbd", (lv.n_gcols,))
        ops.colsum(dLg, B, lv.n_gcols, tmpb.data_ptr())
        dLgi = rt.gemm_input(ws, f"cgc{l}.dlogits_op", dLg, B, lv.n_gcols)
        rt.lin_bwd_w(dLgi, xin, ldw, rt.o(f"cgc{l}.Wbd"), lv.n_gcols, B)          # dWbd (diagonal blocks are the gates' gradients)
        rt.lin_bwd_x(dLgi, ldw, rt.o(f"cgc{l}.Wbd"), lv.n_gcols, dxin, B, accumulate=True)
        for (r0, n, c0, blk, nb, dr, dc, db) in self._gate_runs(lv):
            ops.copy2d_batched(rt.g(f"cgc{l}.Wbd", c0 * ldw + blk * K), dc * ldw + db * K, ldw, rt.g(f"cgc{l}.gW", r0 * K), dr * K, K,
                               nb, n, K, 4)
            ops.copy2d_batched(tmpb.data_ptr() + 4 * c0, dc, n, rt.g(f"cgc{l}.gb", r0), dr, n, nb, 1, n, 4)

    def _dlin_mat(self, ws, B):
        n = self._levels[0].n_gcols
        return ws.mat("cgc0.dlogits", B, n, zero=True).cols(n - 1)

    def _chain_launch(self, ws, l, xin: Mat, B, train, keep_acts) -> Mat:
        """Level l's experts (layer 0 -> layer 1 chained in-kernel) and its gate / wide-linear logits from the same resident X tile
        in ONE launch (cdcmdr_ple_chain_fwd).  The [B, nE*d0] layer-0 activation is written only when a backward will read it
        (keep_acts).  Returns the expert outputs H [B, nE*d1]; the logits land in cgc{l}.logits."""
        rt, lv = self._rt, self._levels[l]
        ex = lv.experts
        d0, d1 = ex.dims
        drop = rt.dropout if train else 0.0
        A0 = ex._act(ws, 0, B) if keep_acts else None
        H = ex._act(ws, 1, B)
        Lg = ws.mat(f"cgc{l}.logits", B, lv.n_gcols)
        names = self._level_names[l]
        rt.ops.ple_chain_fwd(xin, B, lv.K, rt.Wb.data_ptr() + 2 * rt.o(names["W"][0]), rt.w(names["b"][0]),
                             rt.Wb.data_ptr() + 2 * rt.o(names["W"][1]), rt.w(names["b"][1]), ex.G, d0, d1, lv.n_gcols,
                             A0, H, Lg, drop_p=drop, seed_ptr=rt.seed_ptr if drop > 0 else None, salt0=ex.salts[0], salt1=ex.salts[1])
        return H

    def _program_fwd(self, ws, X: Mat, B, train):
        rt = self._rt
        xin = X
        for l, lv in enumerate(self._levels):
            Lg = ws.mat(f"cgc{l}.logits", B, lv.n_gcols)
            if lv.chain and xin.is_bf16:
                H = self._chain_launch(ws, l, xin, B, train, self._keep_acts)
                out = ws.mat(f"cgc{l}.out", B, lv.n_gates * lv.h, rt.act_dtype)
                probs = ws.get(f"cgc{l}.probs", (B, lv.n_gates * lv.max_sel))
                rt.ops.gate_mix_fwd(lv.desc, H, Lg, out, probs, B)
                xin = out
                continue
            H = lv.experts.fwd(ws, xin, B, train)
            if l > 0:
                self._pack_gates(l, lv)
                rt.lin_fwd(xin, lv.n_in * lv.K, rt.o(f"cgc{l}.Wbd"), lv.n_gcols, rt.o(f"cgc{l}.bbd"), Lg, B)
            else:
                for (blk, r0, r1, c0) in lv.gate_groups:
                    rt.lin_fwd(xin.cols(blk * lv.K), lv.K, rt.o(f"cgc{l}.gW", r0 * lv.K), r1 - r0, rt.o(f"cgc{l}.gb", r0),
                               Lg.cols(c0), B)
            out = ws.mat(f"cgc{l}.out", B, lv.n_gates * lv.h, rt.act_dtype)
            probs = ws.get(f"cgc{l}.probs", (B, lv.n_gates * lv.max_sel))
            rt.ops.gate_mix_fwd(lv.desc, H, Lg, out, probs, B)
            xin = out
        logits = self._towers.fwd(ws, xin, B, train)
        n0 = self._levels[0].n_gcols
        lin = ws.mat("cgc0.logits", B, n0).cols(n0 - 1)
        if self._att is not None:                              # ple.py:65-67: one more `other_out` on every tower logit
            self._att.fwd(ws, self._att_x(ws, X, B), B, lin, train)
        return logits, lin

    def _program_bwd(self, ws, X: Mat, B, train, dlogits: Mat):
        rt = self._rt
        T = self.n_tower
        keep = 1.0 / (1.0 - rt.dropout) if (train and rt.dropout > 0) else 1.0
        last_lv = self._levels[-1]
        act = rt.act_dtype
        dcur = ws.mat("towers.dX", B, T * last_lv.h, act)
        self._towers.bwd(ws, ws.mat(f"cgc{self.n_level - 1}.out", B, last_lv.n_gates * last_lv.h, act), dlogits, B, train, dcur)
        for l in reversed(range(self.n_level)):
            lv = self._levels[l]
            nE = lv.experts.G
            H = lv.experts._act(ws, len(lv.experts.dims) - 1, B)
            dH = ws.mat(f"cgc{l}.dH", B, nE * lv.h, act)
            # zero-initialised once: the padding columns between gates are never written and must not feed NaNs into the GEMMs
            dLg = ws.mat(f"cgc{l}.dlogits", B, lv.n_gcols, zero=True)
            probs = ws.get(f"cgc{l}.probs", (B, lv.n_gates * lv.max_sel))
            rt.ops.gate_mix_bwd(lv.desc, H, probs, dcur, dH, keep, dLg, B)
            xin = X if l == 0 else ws.mat(f"cgc{l - 1}.out", B, self._levels[l - 1].n_gates * self._levels[l - 1].h, act)
            # the embedding backward consumes an fp32 gradient; deeper levels hand an activation-dtype gradient to gate_mix_bwd
            dxin = ws.mat(f"cgc{l}.dXin", B, lv.n_in * lv.K, torch.float32 if l == 0 else act)
            if lv.fused_bwd:
                tail = lv.experts.dA0(ws, B).cols(nE * lv.experts.dims[0])
                if rt.bf16:
                    rt.ops.cast_f32_bf16(dLg, tail, B, lv.n_gcols)
                else:
                    rt.ops.add2d(dLg, tail, B, lv.n_gcols, False)
                lv.experts.bwd(ws, xin, dH, B, train, dxin, final_dx=(l == 0 and self._att is None))      # level 0: dxin is the embeddings' gradient
                dcur = dxin
                continue
            lv.experts.bwd(ws, xin, dH, B, train, dxin)
            if l > 0:
                self._gates_bwd_packed(ws, l, lv, xin, dLg, dxin, B)
                dcur = dxin
                continue
            dLgi = rt.gemm_input(ws, f"cgc{l}.dlogits_op", dLg, B, lv.n_gcols)
            for (blk, r0, r1, c0) in lv.gate_groups:
                rt.ops.colsum(dLg.cols(c0), B, r1 - r0, rt.g(f"cgc{l}.gb", r0))
                rt.lin_bwd_w(dLgi.cols(c0), xin.cols(blk * lv.K), lv.K, rt.o(f"cgc{l}.gW", r0 * lv.K), r1 - r0, B)
                rt.lin_bwd_x(dLgi.cols(c0), lv.K, rt.o(f"cgc{l}.gW", r0 * lv.K), r1 - r0, dxin.cols(blk * lv.K), B,
                             accumulate=True)
            dcur = dxin
        if self._att is not None:
            self._att.bwd(ws, self._att_x(ws, X, B), B, self._dlin_mat(ws, B), dcur, train)
        return dcur
