"""CDC model part (reference model/cdc.py:24-119, 343-357) on libcdcmdr.so: a base multi-tower model (PLE / MMoE / STAR)
plus the domain -> cluster routing of its towers.

  forward(x, mode='warmup')                 mean over towers                            cdc.py:99-102
  forward(x, mode='split', domain_i=None)   per-sample tower domain2group[x[:, domain_idx]]   cdc.py:104-107
  forward(x, mode='split', domain_i=d)      fixed tower domain2group_list[d]            cdc.py:108-111
  train_step(...)                           the same three modes fused into the sigmoid+BCE kernel (run.py:635-640)

The clustering itself (update_group: affinity matrices -> causal kernel -> k-means / regrouping, cdc.py:121-341) is
host-side NumPy / SciPy / scikit-learn in the reference and runs once per `update_interval` steps; it is
restated on the host in cdc_group.py so that run.py::train_cdc works unchanged."""
from __future__ import annotations


import numpy as np
import torch

from .cdc_group import Grouping
from .layer import BaseModel
from .mmoe import MMoE
from .ple import PLE


class CDC(BaseModel):
    def __init__(self, feature_dims, embed_dim, n_tower, n_domain, base_model, expert_dims, tower_dims, domain_idx,
                 domain_cnt_weight=None, n_causal_mask=50, use_metric='loss', device='cpu', dropout=0.2, config=None,
                 savefig_folder='', l2_reg_embedding=1e-5, l2_reg_linear=1e-5, l2_reg_dnn=1e-5, l2_reg_cross=1e-5):
        super(BaseModel, self).__init__()          # like the reference: CDC owns no embedding of its own (cdc.py:29)
        self.model_name = 'cdc'
        self.base_model = base_model
        if base_model == 'mmoe':
            self.base_model_instance = MMoE(feature_dims, embed_dim, n_tower, config.mmoe_n_expert, expert_dims, tower_dims,
                                            dropout, config, l2_reg_embedding, l2_reg_linear, l2_reg_dnn, l2_reg_cross,
                                            model_name=self.model_name)
        elif base_model == 'ple':
            self.base_model_instance = PLE(feature_dims, embed_dim, n_tower, config.ple_n_expert_specific,
                                           config.ple_n_expert_shared, expert_dims, tower_dims, dropout, config,
                                           l2_reg_embedding, l2_reg_linear, l2_reg_dnn, l2_reg_cross,
                                           model_name=self.model_name)
        elif base_model == 'star':
            from .star import STAR
            self.base_model_instance = STAR(feature_dims, embed_dim, n_tower, tower_dims, domain_idx, dropout, config,
                                            l2_reg_embedding, l2_reg_linear, l2_reg_dnn, l2_reg_cross, device)
        else:
            raise NotImplementedError(f"CDC base model '{base_model}' is outside the hot path (SURVEY §2: pepnet/epnet)")
        self.device = device
        self.config = config
        self.n_cluster = n_tower
        self.n_causal_mask = n_causal_mask
        self.n_domain = n_domain
        self.domain_idx = domain_idx
        dcw = domain_cnt_weight if domain_cnt_weight is not None else [1.0 / n_domain] * n_domain
        self._dcw = [float(v) for v in dcw]            # the runner's own float64 weights: np.random.choice checks sum(p) == 1 in double
        self.domain_cnt_weight = torch.tensor(dcw, dtype=torch.float32, device=device)
        self.domain2group = torch.zeros(n_domain, dtype=torch.int64, device=device)
        self.domain2group_list = [0] * n_domain
        self.s_group2domain_list = [list(range(n_domain))]
        self.t_group2domain_list = [list(range(n_domain))]
        self.matrix_A = torch.zeros((n_domain + 1, n_domain), dtype=torch.float32, device=device)
        self.matrix_B = torch.zeros((n_domain + self.n_cluster, n_domain), dtype=torch.float32, device=device)
        self.matrix_mask = torch.zeros((n_causal_mask, n_domain), dtype=torch.float32, device=device)
        self.matrix_causal = torch.zeros((n_causal_mask, n_domain), dtype=torch.float32, device=device)
        self.use_metric = use_metric
        self.model_state = None
        self._grouping = Grouping(n_domain, self.n_cluster, dcw, config, use_metric)

    # ---------------------------------------------------------------- grouping state
    def set_groups(self, domain2group_list):
        """Install a domain -> cluster assignment (what update_group computes, cdc.py:236-238)."""
        d2g = [int(g) for g in domain2group_list]
        if len(d2g) != self.n_domain or min(d2g) < 0 or max(d2g) >= self.n_cluster:
            raise ValueError("domain2group must map every domain to a cluster in [0, n_cluster)")
        self.domain2group_list = d2g
        self._install_domain2group(d2g)
        self._grouping.domain2group_list = list(d2g)
        # the source / target domain sets that go with an assignment installed by hand: every cluster trains on its own members
        # (what update_group leaves behind when no cluster borrows domains; run.py:580-592 indexes these lists by cluster id)
        members = [[d for d in range(self.n_domain) if d2g[d] == g] for g in range(self.n_cluster)]
        self.s_group2domain_list = [list(m) for m in members]
        self.t_group2domain_list = [list(m) for m in members]
        self._grouping.s_group2domain_list = [list(m) for m in members]
        self._grouping.t_group2domain_list = [list(m) for m in members]

    def _install_domain2group(self, d2g):
        """In place: a captured CUDA graph (mode 'gather') addresses this tensor; the generation counter tells GraphedTrainStep
        that a baked tower column (mode 'col') is stale."""
        self.domain2group.copy_(torch.tensor(d2g, dtype=torch.int64))
        self._group_gen = getattr(self, "_group_gen", 0) + 1

    def _apply(self, fn, recurse=True):
        out = torch.nn.Module._apply(self, fn)
        dev = next(self.base_model_instance.parameters()).device
        for k in ("domain2group", "domain_cnt_weight", "matrix_A", "matrix_B", "matrix_mask", "matrix_causal"):
            setattr(self, k, getattr(self, k).to(dev))            # plain attributes in the reference (cdc.py:66-76)
        return out

    # ---------------------------------------------------------------- forward (cdc.py:95-111)
    def forward(self, x, mode='split', domain_i=None):
        base = self.base_model_instance
        y_cat = base._call(x)
        if mode == 'warmup':
            return torch.mean(y_cat, dim=1)
        if mode == 'split':
            if domain_i is None:
                return y_cat.gather(1, self._groups_of(x).unsqueeze(1))
            return y_cat[:, self.domain2group_list[domain_i]]
        raise ValueError(f"unknown CDC mode {mode!r}")

    def _groups_of(self, x):
        base = self.base_model_instance
        rt = base._rt
        B = x.shape[0]
        groups = torch.empty(B, dtype=torch.int64, device=x.device)
        d2g = self.domain2group.to(x.device)
        rt.ops.lib.domain_to_group(x.data_ptr(), B, x.shape[1], self.domain_idx, d2g.data_ptr(), self.n_domain,
                                   groups.data_ptr(), rt.ops.stream)
        return groups

    def train_step(self, x, y, optimizer, mode='split', domain_i=None, x_next=None, prefetched=False):
        """run.py:616-622 (warmup) / 635-640 (split), fused.  x_next / prefetched: BaseModel.train_step (exchange pipelining)."""
        base = self.base_model_instance
        pf = dict(x_next=x_next, prefetched=prefetched)
        if mode == 'warmup':
            return base.train_step(x, y, optimizer, mode="mean", **pf)
        if mode != 'split':
            raise ValueError(f"unknown CDC mode {mode!r}")
        if domain_i is None:
            base._check_device(x)
            return base.train_step(x, y, optimizer, mode="gather", sel=self._groups_of(x.contiguous()), **pf)
        return base.train_step(x, y, optimizer, mode="col", col=self.domain2group_list[domain_i], **pf)

    step_losses = staticmethod(BaseModel.step_losses)

    def get_regularization_loss(self, device=None):
        return self.base_model_instance.get_regularization_loss(device)

    def get_matrix_metric(self, preds, targets):
        """cdc.py:113-119.  'auc' is scikit-learn on the host upstream and stays there (one affinity entry per call)."""
        if self.use_metric == 'loss':
            return torch.nn.functional.binary_cross_entropy(preds, targets).detach()
        if self.use_metric == 'auc':
            from sklearn.metrics import roc_auc_score
            return roc_auc_score(targets.detach().cpu().numpy(), preds.detach().cpu().numpy())
        raise ValueError(f"unknown use_metric {self.use_metric!r}")

    def probe_all_domains(self, batches, max_rows=None):
        """One row of the affinity matrices (run.py:551-560 `cdc_test_all_domain`): the BCE of every domain's batch under the current
        weights.  The reference runs n_domain forwards of `model(X_d, mode='split', domain_i=d)`, each followed by
        `get_matrix_metric`; here the batches are concatenated, evaluated in ONE pass with the tower of every row's domain
        (`domain2group[x[:, domain_idx]]` - the same tower `domain_i=d` selects for rows of domain d) and reduced per domain on the
        device (cdcmdr_bce_segments).  batches: sequence of (X_d int32 [B_d, F], y_d [B_d]) for d = 0..n_domain-1.
        max_rows caps the rows of one evaluation (consecutive domains are grouped up to that many rows; the workspace of the
        probe is proportional to it - 30 x 65 536 rows hold ~25 GB of activations at the C4 shape); results are identical.
        Returns a float32 tensor [len(batches)] on the model's device; the caller assigns it to matrix_mask / matrix_A / matrix_B."""
        if self.use_metric != 'loss':
            return self._probe_per_domain(batches)                # 'auc': scikit-learn on the host per domain, as upstream
        if max_rows is None:
            max_rows = self.PROBE_ROWS
        n_seg = len(batches)
        dev = batches[0][0].device
        out = torch.empty(n_seg, dtype=torch.float32, device=dev)
        was_training = self.training
        self.eval()
        try:
            first = 0
            while first < n_seg:
                last, rows = first, 0
                while last < n_seg and (last == first or max_rows is None or rows + int(batches[last][0].shape[0]) <= max_rows):
                    rows += int(batches[last][0].shape[0])
                    last += 1
                self._probe_chunk(batches[first:last], out[first:last])
                first = last
        finally:
            self.train(was_training)
        return out

    def _probe_per_domain(self, batches):
        """run.py:550-558 as written: one evaluation per domain with `domain_i=d`, each followed by get_matrix_metric."""
        was_training = self.training
        self.eval()
        try:
            row = []
            with torch.no_grad():
                for d, (x, y) in enumerate(batches):
                    pred = self.forward(x, mode='split', domain_i=d)
                    row.append(float(self.get_matrix_metric(pred.reshape(-1), y.reshape(-1).float())))
        finally:
            self.train(was_training)
        return torch.tensor(row, dtype=torch.float32, device=batches[0][0].device)

    def _probe_chunk(self, batches, out):
        rt = self.base_model_instance._rt
        xs = [b[0] for b in batches]
        ys = [b[1].reshape(-1) for b in batches]
        if sum(int(xb.shape[0]) for xb in xs) == 0:
            out.fill_(float("nan"))                                                  # torch's mean of an empty tensor (cdc.py:116-119)
            return
        x = torch.cat(xs, dim=0).contiguous()
        y = torch.cat(ys, dim=0)
        y = (y if y.dtype in (torch.int16, torch.float32) else y.to(torch.float32)).contiguous()
        bounds = [0]
        for xb in xs:
            bounds.append(bounds[-1] + int(xb.shape[0]))
        seg = torch.tensor(bounds, dtype=torch.int64, device=x.device)
        with torch.no_grad():
            pred = self.forward(x, mode='split').reshape(-1).contiguous()            # (sum B_d,) fp32 probabilities
        n_seg = len(xs)
        lib = rt.ops.lib
        sc = rt.ops.scratch("bce_segments", lib.bce_segments_scratch_bytes(n_seg))
        lib.bce_segments(pred.data_ptr(), 1, y.data_ptr(), 1 if y.dtype == torch.float32 else 0, seg.data_ptr(), n_seg,
                         out.data_ptr(), sc.data_ptr(), rt.ops.stream)

    # ---------------------------------------------------------------- the affinity-matrix probing loop (run.py:528-594)
    # default cap on the rows of one batched probe evaluation: the activations of 30 x 65 536 rows are ~25 GB, which together with
    # the training workspaces of the probe steps overflows the workspace budget (Runtime.WS_MAX_BYTES) and turns every probe into
    # a re-allocation; 2^19 rows keep one probe evaluation at ~7 GB with identical numbers
    PROBE_ROWS = 1 << 19

    def update_matrix_cdc(self, get_domain_data, optimizer, update_matrix_step, rng=None, probe_rows=None):
        """`Run.update_matrix_cdc` (run.py:528-594) as one call: snapshot -> for every probe {k fused training steps on a domain
        (multi)set -> one batched evaluation of all domains -> restore} -> update_group().  Returns the new domain2group_list
        (the caller stores it: `self.domain2group_list = model.update_matrix_cdc(self.get_domain_data, optimizer, k)`).

        get_domain_data(d) is the runner's own batch provider (run.py:499-526): an int yields that domain's next (X, y), a list
        yields the concatenation over the (shuffled in place) list.  Draws, shuffles and fetches happen in the reference's order
        from NumPy's global RNG (or `rng`, a RandomState), so a seeded run probes the same multisets on the same batches.
        probe_rows caps the rows of one batched evaluation (probe_all_domains(max_rows=...)).
        Faithful to two upstream quirks, because they change the numbers: (1) cdc_test_all_domain leaves the model in eval mode
        and only the matrix-A loop switches back, so the mask probes after the first one and every matrix-B probe TRAIN in eval
        mode (BatchNorm on running statistics, dropout off), and the model is still in eval mode on return; (2) for the rows of
        matrix B past n_domain the training "set" is `domain2group_list[d_i - n_domain]` - an int, i.e. the single domain whose
        index equals that group id (run.py:587-588)."""
        rs = rng if rng is not None else np.random
        n_domain, n_cluster = self.n_domain, self.n_cluster
        p = self._dcw

        def train_with(domains, k):                                  # cdc_train_update_with_domain (run.py:529-548), mode 'split'
            if isinstance(domains, (int, np.integer)):
                seq = [domains] * k
            else:
                tmp = list(domains) * k
                seq = [tmp[i:i + 7] for i in range(0, len(tmp), 7)]
            for one in seq:
                X, y = get_domain_data(one)
                self.train_step(X, y, optimizer, mode='split', domain_i=one if isinstance(one, (int, np.integer)) else None)

        def probe():                                                 # cdc_test_all_domain (run.py:550-558)
            self.eval()
            return self.probe_all_domains([get_domain_data(d) for d in range(n_domain)],
                                          max_rows=probe_rows if probe_rows is not None else self.PROBE_ROWS)

        self.save_model_state()
        for line_i in range(self.n_causal_mask):                     # treatment rows (run.py:563-569)
            size = rs.randint(5, n_domain)
            domains = rs.choice(range(n_domain), p=p, size=size)
            train_with(domains, update_matrix_step)
            self.matrix_mask[line_i] = probe()
            self.load_model_state()
        self.matrix_A[n_domain] = probe()                            # matrix A (run.py:572-577)
        for d_i in range(n_domain):
            self.train()
            train_with(d_i, update_matrix_step)
            self.matrix_A[d_i] = probe()
            self.load_model_state()
        n_rows = n_domain + n_cluster if max(self.domain2group_list) > 0 else n_domain + 1      # matrix B (run.py:580-592)
        for d_i in range(n_rows):
            if d_i >= n_domain:
                domains = self.domain2group_list[d_i - n_domain]
            else:
                domains = [d for d in self.s_group2domain_list[self.domain2group_list[d_i]] if d != d_i]
            train_with(domains, update_matrix_step)
            self.matrix_B[d_i] = probe()
            self.load_model_state()
        return self.update_group()

    # ---------------------------------------------------------------- snapshot / restore (cdc.py:343-354)
    def save_model_state(self):
        """cdc.py:343-351: upstream deep-copies every `base_model_instance.*` entry of the state_dict (parameters and buffers, not the
        optimizer's moments).  Same content here, kept ON the device as three flat copies - the dense parameter arena, the BatchNorm
        buffer arena and the embedding table - plus the handful of integer `num_batches_tracked` counters: the probing loop restores
        115 times per update (run.py:560-592), and a state_dict round trip per restore was a quarter of its time.  (Every row of the
        table moves every step under the reference's dense Adam + full-table L2 - SURVEY G6 - so the table copy is the exact
        snapshot; it is one 64 MB device copy at the C4 shape.)"""
        base = self.base_model_instance
        rt = base._rt
        table = base.embedding.embedding_dict.weight
        ints = [b for n, b in base.named_buffers() if not b.dtype.is_floating_point and b.device == rt.device and n != "embedding.offsets_dev"]
        self.model_state = dict(W=rt.W.clone(), Bf=rt.Bf.clone(), table=table.detach().clone(), ints=ints,
                                int_vals=[b.clone() for b in ints], rt=rt)

    def load_model_state(self):
        st = self.model_state
        base = self.base_model_instance
        rt = base._rt
        if st is None:
            raise RuntimeError("cdcmdr: load_model_state() before save_model_state()")
        if st["rt"] is not rt:
            raise RuntimeError("cdcmdr: the model moved to another device since save_model_state()")
        with torch.no_grad():
            rt.W.copy_(st["W"])
            rt.Bf.copy_(st["Bf"])
            base.embedding.embedding_dict.weight.copy_(st["table"])
            if st["ints"]:
                torch._foreach_copy_(st["ints"], st["int_vals"])

    def update_group(self, mode='iterative'):
        """cdc.py:121-238: turn the probed affinity matrices (matrix_A / matrix_B / matrix_mask, filled row by row by
        run.py::update_matrix_cdc) into a new domain -> cluster assignment.  Host-side (cdc_group.py)."""
        g = self._grouping
        dev = self.domain2group.device
        out = g.update(self.matrix_A.detach().cpu().numpy(), self.matrix_B.detach().cpu().numpy(),
                       self.matrix_mask.detach().cpu().numpy(), mode=mode)
        self.matrix_A = torch.from_numpy(out["A"]).to(dev)
        self.matrix_B = torch.from_numpy(out["B"]).to(dev)
        self.matrix_mask = torch.from_numpy(out["mask"]).to(dev)
        self.matrix_causal = torch.from_numpy(out["causal"]).to(dev)
        self.domain2group_list = list(g.domain2group_list)
        self._install_domain2group(self.domain2group_list)
        self.s_group2domain_list = [list(v) for v in g.s_group2domain_list]
        self.t_group2domain_list = [list(v) for v in g.t_group2domain_list]
        return self.domain2group_list

    @property
    def call_update_group(self):
        return self._grouping.call_update_group

    @property
    def p_weight(self):
        return self._grouping.p_weight
