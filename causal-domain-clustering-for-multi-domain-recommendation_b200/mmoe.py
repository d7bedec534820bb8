"""MMoE (reference model/mmoe.py:10-74) on libcdcmdr.so: all experts' first layers as one concatenated-N GEMM, deeper
layers grouped, BatchNorm+ReLU fused per layer, gate softmax + expert mixing as one kernel (mmoe.py:56-60)."""
from __future__ import annotations

import torch
from torch import nn

from .core import Mat
from .layer import BaseModel, MultiLayerPerceptron, mlp_group_names, precision_of
from .runtime import MlpGroup


class MMoE(BaseModel):
    def __init__(self, feature_dims, embed_dim, n_tower, n_expert, expert_dims, tower_dims, dropout=0.2, config=None,
                 l2_reg_embedding=1e-5, l2_reg_linear=1e-5, l2_reg_dnn=1e-5, l2_reg_cross=1e-5, model_name='mmoe'):
        super().__init__(feature_dims, embed_dim, l2_reg_embedding=l2_reg_embedding, l2_reg_linear=l2_reg_linear)
        self.config = config
        self.model_name = model_name
        self.n_tower = n_tower
        self.n_out = n_tower
        self.n_expert = n_expert
        if getattr(config, 'use_dcn', False):
            raise NotImplementedError("use_dcn=True is broken upstream (SURVEY G5) and not part of the hot path")
        if getattr(config, 'use_atten', False):
            self.build_atten(config, dropout)                  # mmoe.py: build_atten before the experts, like upstream
        self.expert_dims, self.tower_dims = tuple(expert_dims), tuple(tower_dims)
        self.experts = nn.ModuleList(MultiLayerPerceptron(self.embed_output_dim, expert_dims, dropout, output_layer=False)
                                     for _ in range(n_expert))
        self.gates = nn.ModuleList([nn.Sequential(nn.Linear(self.embed_output_dim, n_expert), nn.Softmax(dim=1))
                                    for _ in range(n_tower)])
        self.towers, self.towers_linear, self.output_layers = self.build_tower_output(n_tower, expert_dims[-1], tower_dims,
                                                                                      dropout)
        self.add_regularization_weight(self.reg_filter("experts"), l2=l2_reg_dnn)
        self.add_regularization_weight(self.reg_filter("towers"), l2=l2_reg_dnn)

        enames, eblk, ebufs = mlp_group_names([f"experts.{i}" for i in range(n_expert)], self.experts[0], "experts")
        gw = [f"gates.{t}.0.weight" for t in range(n_tower)] + ["linear.fc.weight"]
        gb = [f"gates.{t}.0.bias" for t in range(n_tower)] + ["linear.fc.bias"]
        tnames, tblk, tbufs = mlp_group_names([f"towers.{t}" for t in range(n_tower)], self.towers[0], "towers")
        self._expert_names, self._tower_names = enames, tnames
        blocks = [eblk[0], ("gates.W", gw), eblk[1], ("gates.b", gb)] + eblk[2:] + tblk
        self._finalize(blocks, ebufs + tbufs, precision=precision_of(config), dropout=dropout)

    def _on_runtime_built(self):
        rt = self._rt
        T, nE, D = self.n_tower, self.n_expert, self.embed_output_dim
        self._experts = MlpGroup(rt, "experts", nE, D, self.expert_dims, self._expert_names, bn=True, out_layer=False,
                                 in_groups=[(0, 0, nE)])
        self._towers = MlpGroup(rt, "towers", T, self.expert_dims[-1], self.tower_dims, self._tower_names, bn=True,
                                out_layer=True, in_groups=None)
        self._n_gcols = T * nE + 1
        col = [t * nE for t in range(T)]
        self._desc_t = torch.tensor(col + [nE] * T + list(range(nE)) * T, dtype=torch.int32, device=rt.device)
        self._desc = rt.ops.mix_desc(T, nE, self.expert_dims[-1], nE, self._desc_t, n_pairs=T * nE)

    def _dlin_mat(self, ws, B):
        return ws.mat("gates.dlogits", B, self._n_gcols).cols(self._n_gcols - 1)

    def _program_fwd(self, ws, X: Mat, B, train):
        rt = self._rt
        T, nE, D, h = self.n_tower, self.n_expert, self.embed_output_dim, self.expert_dims[-1]
        H = self._experts.fwd(ws, X, B, train)
        Lg = ws.mat("gates.logits", B, self._n_gcols)
        rt.lin_fwd(X, D, rt.o("gates.W"), self._n_gcols, rt.o("gates.b"), Lg, B)
        out = ws.mat("mix.out", B, T * h, rt.act_dtype)
        probs = ws.get("mix.probs", (B, T * nE))
        rt.ops.gate_mix_fwd(self._desc, H, Lg, out, probs, B)
        logits = self._towers.fwd(ws, out, B, train)
        lin = Lg.cols(self._n_gcols - 1)
        if self._att is not None:
            self._att.fwd(ws, self._att_x(ws, X, B), B, lin, train)
        return logits, lin

    def _program_bwd(self, ws, X: Mat, B, train, dlogits: Mat):
        rt = self._rt
        T, nE, D, h = self.n_tower, self.n_expert, self.embed_output_dim, self.expert_dims[-1]
        act = rt.act_dtype
        out = ws.mat("mix.out", B, T * h, act)
        dout = ws.mat("mix.dout", B, T * h, act)
        self._towers.bwd(ws, out, dlogits, B, train, dout)
        H = self._experts._act(ws, len(self.expert_dims) - 1, B)
        dH = ws.mat("mix.dH", B, nE * h, act)
        dLg = ws.mat("gates.dlogits", B, self._n_gcols)
        probs = ws.get("mix.probs", (B, T * nE))
        if B == 1:      # BatchNorm skipped (layer.py:202-204): the ReLU/dropout mask is applied here instead
            keep = 1.0 / (1.0 - rt.dropout) if (train and rt.dropout > 0) else 1.0
            rt.ops.gate_mix_bwd(self._desc, H, probs, dout, dH, keep, dLg, B)
        else:
            rt.ops.gate_mix_bwd(self._desc, H, probs, dout, dH, 0.0, dLg, B)
        dX = ws.mat("dX", B, D)
        self._experts.bwd(ws, X, dH, B, train, dX)
        rt.ops.colsum(dLg, B, self._n_gcols, rt.g("gates.b"))
        dLgi = rt.gemm_input(ws, "gates.dlogits_op", dLg, B, self._n_gcols)
        rt.lin_bwd_w(dLgi, X, D, rt.o("gates.W"), self._n_gcols, B)
        rt.lin_bwd_x(dLgi, D, rt.o("gates.W"), self._n_gcols, dX, B, accumulate=True)
        if self._att is not None:
            self._att.bwd(ws, self._att_x(ws, X, B), B, self._dlin_mat(ws, B), dX, train)
        return dX
