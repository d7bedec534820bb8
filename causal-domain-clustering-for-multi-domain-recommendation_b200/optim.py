"""Drop-in for `torch.optim.Adam(model.parameters(), lr, betas, eps, weight_decay)` (run.py:720-721) that updates
the flat parameter arena and the embedding table with the fused CUDA kernels of libcdcmdr.so.

Two ways in:
  * `model.train_step(x, y, optimizer, ...)`  - the fused step; this class only supplies hyper-parameters and the
    device-resident step state (optimizer.tick).
  * `loss.backward(); optimizer.step()`        - run.py's loop unchanged: step() consumes the .grad tensors autograd
    filled (they already contain the regulariser's gradient) and applies the same Adam kernels.
The arithmetic follows torch's _single_tensor_adam: L2-style weight decay added to the gradient,
m <- m + (1-b1)(g-m), v <- b2 v + (1-b2) g^2, p <- p - lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps).
"""
from __future__ import annotations

import torch


class Adam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, seed=2000):
        self.params = list(params)
        self.defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        self.param_groups = [dict(params=self.params, **self.defaults)]
        self.seed = int(seed)
        self.model = None
        self.steps = 0

    # ---------------------------------------------------------------- plumbing
    def attach(self, model):
        if self.model is None:
            self.model = model
        elif self.model is not model:
            raise RuntimeError("cdcmdr Adam is bound to one model")

    def _find_model(self):
        if self.model is None:
            raise RuntimeError("cdcmdr Adam: call optimizer.attach(model) (or use model.train_step) before step()")
        return self.model

    def tick(self, rt):
        g = self.param_groups[0]
        rt.ensure_opt_state()
        rt.ops.step_tick(rt.step_state, g["lr"], g["betas"], g["eps"], g["weight_decay"], self.seed)
        self.steps += 1

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    # ---------------------------------------------------------------- autograd path
    @torch.no_grad()
    def step(self):
        model = self._find_model()
        rt = model._rt
        self.tick(rt)
        present = torch.zeros_like(rt.present)
        table = model.embedding.embedding_dict.weight
        for name, p in model.named_parameters():
            if p is table or p.grad is None:
                continue
            o = rt.params.off[name]
            rt.G[o:o + p.numel()].copy_(p.grad.reshape(-1))
            present[o:o + p.numel()] = 1
        rt.ops.adam_dense(rt.W, rt.G, rt.M, rt.V, None, present, rt.W.numel(), rt.step_state)
        if table.grad is not None:
            if model._table_state is None:
                model._table_state = (torch.zeros_like(table), torch.zeros_like(table))
            m, v = model._table_state
            rt.ops.adam_dense(table, table.grad.contiguous(), m, v, None, None, table.numel(), rt.step_state)

    # ---------------------------------------------------------------- checkpointing (run.py:447-459)
    def state_dict(self):
        model = self.model
        out = dict(param_groups=[{k: v for k, v in g.items() if k != "params"} for g in self.param_groups],
                   steps=self.steps, seed=self.seed)
        if model is not None and model._rt.M is not None:
            out["exp_avg"] = model._rt.M.clone()
            out["exp_avg_sq"] = model._rt.V.clone()
        if model is not None and model._table_state is not None:
            out["table_exp_avg"], out["table_exp_avg_sq"] = (t.clone() for t in model._table_state)
        dp = getattr(model._rt, "dp", None) if model is not None else None
        if dp is not None and dp.shard and dp._moments is not None:
            # row-sharded table: the moments of the embedding rows live with their owner; each rank checkpoints its own row range
            out["table_shard_rows"] = (dp.row0, dp.row1)
            out["table_shard_exp_avg"], out["table_shard_exp_avg_sq"] = (t.clone() for t in dp._moments)
        return out

    def load_state_dict(self, sd):
        model = self._find_model()
        rt = model._rt
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update(s)
        self.steps, self.seed = int(sd["steps"]), int(sd["seed"])
        rt.ensure_opt_state()
        if "exp_avg" in sd:
            rt.M.copy_(sd["exp_avg"]); rt.V.copy_(sd["exp_avg_sq"])
        if "table_exp_avg" in sd:
            model._table_state = (sd["table_exp_avg"].to(rt.device).clone(), sd["table_exp_avg_sq"].to(rt.device).clone())
        if "table_shard_rows" in sd:
            dp = getattr(rt, "dp", None)
            if dp is None or not dp.shard or tuple(sd["table_shard_rows"]) != (dp.row0, dp.row1):
                raise RuntimeError("cdcmdr Adam: the checkpoint holds the embedding moments of row range "
                                   f"{tuple(sd['table_shard_rows'])}; load it on the rank that owns those rows (same world size)")
            dp._moments = (sd["table_shard_exp_avg"].to(rt.device).clone(), sd["table_shard_exp_avg_sq"].to(rt.device).clone())
        rt.ops.step_state_set(rt.step_state, self.steps)
