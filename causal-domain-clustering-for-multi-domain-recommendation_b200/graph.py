"""CUDA-graph capture of the fused training step: the whole step (about a hundred launches of libcdcmdr.so kernels,
no host synchronisation, all state device-resident) is recorded once per batch size and replayed with one launch."""
from __future__ import annotations

import os

import torch


class GraphedTrainStep:
    """Captures `model.train_step(x, y, optimizer, **kw)` over static input buffers.

    usage:  step = GraphedTrainStep(model, optimizer, B, F, mode='split', domain_i=3)
            step.x.copy_(batch_x, non_blocking=True); step.y.copy_(batch_y, non_blocking=True); out = step()

    prefetch=True (data-parallel replicas with the field-sharded table): the recorded step skips its own embedding exchange and
    instead issues the exchange of `step.x_next` - the NEXT batch's indices, which the caller fills before every replay - behind
    its table update (BaseModel.train_step).  `step.prime()` runs the exchange of `step.x` once, eagerly; it is needed before the
    first replay (capture() does it) and after anything else ran a forward at this batch size.
    """

    def __init__(self, model, optimizer, B, F, y_dtype=torch.int16, warmup=2, prefetch=False, **kw):
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("CUDA graphs need a CUDA device")
        self.model, self.optimizer, self.kw = model, optimizer, kw
        self.x = torch.zeros(B, F, dtype=torch.int32, device=dev)
        self.y = torch.zeros(B, dtype=y_dtype, device=dev)
        self.prefetch = bool(prefetch)
        self.x_next = torch.zeros(B, F, dtype=torch.int32, device=dev) if self.prefetch else None
        self.graph = None
        self.out = None
        self.warmup = warmup
        self.launches_per_step = None
        self._gen = None

    def _routing_generation(self):
        """CDC bakes its routing into the recorded step (mode 'col': the tower column of `domain_i` is a kernel argument); every
        regrouping (CDC.update_group / set_groups, run.py:594 - once per update_interval steps) bumps this counter."""
        return getattr(self.model, "_group_gen", 0)

    def capture(self):
        """Call after x / y hold a valid batch (the warm-up steps are REAL optimizer steps on that batch)."""
        rt = self.model_base()._rt
        lib = rt.ops.lib
        rt.pin_ws(self.x.shape[0])                      # the graph holds this workspace's addresses: never evicted
        if self.graph is None:                          # a re-capture (routing changed) must not take extra optimizer steps
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(self.warmup):
                    self.model.train_step(self.x, self.y, self.optimizer, **self.kw)
            torch.cuda.current_stream().wait_stream(side)
        if self.prefetch:
            self.prime()                                # X of the first replay; the recorded step consumes it and prefetches x_next
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        n0 = lib.launch_count()
        step_kw = dict(self.kw, x_next=self.x_next, prefetched=True) if self.prefetch else self.kw
        # The main branch is captured on a HIGH-priority stream (the side branch's stream has the default, lowest priority): kernel
        # nodes inherit it, so whenever both branches have CTAs waiting for an SM the model program's go first - the side branch fills
        # what is left (it used to hold the split-K reduce behind the table sweep for ~30 us at the end of the step).
        cap = torch.cuda.Stream(device=self.x.device, priority=-1) if os.environ.get("CDCMDR_GRAPH_PRIORITY", "1") != "0" else None
        with torch.cuda.graph(self.graph, stream=cap):
            self.out = self.model.train_step(self.x, self.y, self.optimizer, **step_kw)
        self.launches_per_step = int(lib.launch_count() - n0)
        self.optimizer.steps -= 1          # capture records the step but does not execute it
        self._gen = self._routing_generation()
        return self

    def prime(self):
        """Eager exchange of `self.x` (indices to the owners, rows back into X): what the previous step's prefetch would have left."""
        base = self.model_base()
        rt = base._rt
        B = self.x.shape[0]
        base._gather(rt.ws(B), self.x, B, phase="exchange")

    def model_base(self):
        return getattr(self.model, "base_model_instance", self.model)

    def __call__(self):
        if self.graph is None or self._gen != self._routing_generation():
            self.capture()                 # first use, or CDC regrouped since the capture: the recorded tower column is stale
        self.graph.replay()
        self.optimizer.steps += 1
        return self.out
