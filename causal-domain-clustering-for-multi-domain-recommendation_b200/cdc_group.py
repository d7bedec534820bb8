"""Host side of CDC's causal domain clustering (reference model/cdc.py:121-341, 359-393).

This is small float32 / float64 matrix logic over n_domain x n_domain affinities (30..50 domains) that the reference runs once
per `update_interval` training steps on the host; it stays on the host here too (NumPy / SciPy / scikit-learn, as upstream).
It is restated - not accelerated - so that `CDC.update_group()` works and `run.py::train_cdc` runs unchanged against this
package.  The xlsx / png dumps of the matrices (cdc.py:395-426) are an I/O side effect outside the hot path and are not produced.

Pinned against the unmodified reference in tests/test_cdc_group.py (same matrices, same NumPy seed for the unseeded KMeans).
"""
from __future__ import annotations

import copy

import numpy as np
import torch

F32 = np.float32


def calc_causal_matrix(X, alpha=None):
    """Distance-covariance dependence kernel between the columns of X (cdc.py:364-393): per feature j the doubly-centred,
    mean-normalised |x_i - x_k| matrix Z_j; gamma = (F^T F)^2 - 2 <Z, Z>_thresh + |thresh|; cosine-normalised, clipped at 1."""
    from scipy.spatial.distance import pdist, squareform
    X = np.asarray(X)
    n, nf = X.shape
    thresh = np.eye(nf)
    if alpha is not None:
        from scipy.stats import chi2
        thresh[thresh == 0] = chi2(1).ppf(1 - alpha) / n
        thresh[thresh == 1] = 0
    Z = np.zeros((nf, n, n))
    for j in range(nf):
        D = squareform(pdist(X[:, j].reshape(-1, 1), "cityblock"))
        Z[j] = ((D - D.mean(0) - D.mean(1).reshape(-1, 1)) / D.mean()) + 1
    Fm = Z.reshape(nf * n, n)
    left = np.tensordot(Z, thresh, axes=([0], [0]))
    left_right = np.tensordot(left, Z, axes=([2, 1], [0, 1]))
    gamma = (Fm.T @ Fm) ** 2 - 2 * left_right + np.linalg.norm(thresh)
    diag = np.diag(gamma)
    kappa = gamma / np.sqrt(np.outer(diag, diag))
    kappa[kappa > 1] = 1
    return kappa


class Grouping:
    """The mutable clustering state of one CDC model (the attributes cdc.py keeps on the module) and its update rule."""

    def __init__(self, n_domain, n_cluster, domain_cnt_weight, config, use_metric="loss"):
        self.n_domain, self.n_cluster = n_domain, n_cluster
        self._all_domains = torch.arange(n_domain, dtype=torch.int64)
        self.w = np.asarray(domain_cnt_weight, dtype=F32)
        self.w_t = torch.tensor([float(v) for v in domain_cnt_weight], dtype=torch.float32)      # cdc.py:62
        self.affinity_func = getattr(config, "affinity_func", "minus")
        self.p_weight0 = getattr(config, "p_weight", 0.1)
        self.p_weight = self.p_weight0
        self.p_weight_method = getattr(config, "p_weight_method", "linear_decay")
        self.p_weight_exp_decay = getattr(config, "p_weight_exp_decay", 0.9)
        self.old_matrix_weight = getattr(config, "old_matrix_weight", 0.0)
        self.domain2group_list = [0] * n_domain
        self.s_group2domain_list = [list(range(n_domain))]
        self.t_group2domain_list = [list(range(n_domain))]
        self.initial_s_group2domain_list = None
        self.call_update_group = 0
        self.old_A = self.old_B = self.old_mask = None
        if (use_metric == "loss") ^ (self.affinity_func == "divide"):          # cdc.py:88-93
            self.default_metric_value, self.max_better = F32(1e6), False
        else:
            self.default_metric_value, self.max_better = F32(-1e6), True
        self.A = self.B = self.causal = None
        self.A_t = self.B_t = self.causal_t = None
        self._memo = {}

    # ---------------------------------------------------------------- cdc.py:296-306
    def _update_p_weight(self):
        if self.p_weight > 1e-10:
            if self.p_weight_method == "linear_decay":
                self.p_weight = self.p_weight0 / self.call_update_group
            elif self.p_weight_method == "quadratic_decay":
                self.p_weight = self.p_weight0 / (self.call_update_group ** 2)
            elif self.p_weight_method == "exponential_decay":
                self.p_weight = self.p_weight * self.p_weight_exp_decay

    # The affinity arithmetic below runs on torch CPU float32 tensors, operation for operation in upstream's order: the regrouping
    # compares sums over the SAME domains taken in different orders (two source groups that both grew to contain every domain),
    # so its decisions hang on the last bit of a float32 reduction - only torch's own reduction order reproduces them (found by
    # differential fuzzing against the reference: 3 of 480 calls disagreed with NumPy sums, 0 with these).
    # ---------------------------------------------------------------- cdc.py:321-341
    # Within ONE update() the matrices, p_weight and the initial source groups are fixed, so the three helpers below are pure
    # functions of their (small, hashable) arguments - and the regrouping asks the same questions again and again (30-domain case:
    # 76 source-group growths over 30 distinct target groups, 1 004 metrics over 410 distinct pairs).  update() clears `_memo`;
    # a hit returns the very tensor the first call computed, so nothing about the arithmetic or its order changes.
    def _lambda(self, group, domain=None):
        key = ("lam", tuple(group), None if domain is None else tuple(domain))
        hit = self._memo.get(key)
        if hit is None:
            hit = self._memo[key] = self._lambda_uncached(group, domain)
        return hit

    def _lambda_uncached(self, group, domain=None):
        """lambda_d = clamp(0.5 (|G|-1) sum_{g in G} dist[g, d] / (sum_{GxG} dist - sum_{g in G} dist[g, d]), 0, 1)"""
        g = torch.as_tensor(list(group), dtype=torch.int64)
        dom = self._all_domains if domain is None else torch.as_tensor(list(domain), dtype=torch.int64)
        total = torch.sum(self.causal_t[g[:, None], g[None, :]])
        related = torch.sum(self.causal_t[g[:, None], dom[None, :]], dim=0)
        vals = (len(g) - 1) * related / (total - related) * 0.5
        return torch.clamp(vals, min=0, max=1)                  # NaN (0/0) stays NaN

    def _lambda_candidates(self, s_group, cands, domain):
        """_lambda(s_group + [d], domain) for every candidate d at once -> [len(cands), len(domain)].  One gather and one reduction
        per term instead of one pair per candidate; every slice reduces the same elements in the same order as the single call
        (checked bit for bit against it: torch reduces each [m, m] / [m, n] slice of the stacked tensor like the lone matrix)."""
        idx = torch.as_tensor([list(s_group) + [d] for d in cands], dtype=torch.int64)          # [k, m+1]
        dom = torch.as_tensor(list(domain), dtype=torch.int64)
        k, m1 = idx.shape
        rows = self.causal_t.index_select(0, idx.reshape(-1)).view(k, m1, self.n_domain)        # rows of every candidate group
        total = torch.gather(rows, 2, idx[:, None, :].expand(k, m1, m1)).sum(dim=(1, 2))        # [k]
        related = rows.index_select(2, dom).sum(dim=1)                                          # [k, n]
        vals = (idx.shape[1] - 1) * related / (total[:, None] - related) * 0.5
        return torch.clamp(vals, min=0, max=1)

    # ---------------------------------------------------------------- cdc.py:314-319
    def _centers(self, group, center_num=1):
        k = min(center_num, len(group))
        order = torch.topk(self._lambda(group, group), k=k, largest=False)[1]
        return [group[int(i)] for i in order]

    # ---------------------------------------------------------------- cdc.py:308-312
    def _metric_in_source_group(self, target, s_group):
        key = ("met", int(target), tuple(s_group))
        hit = self._memo.get(key)
        if hit is None:
            lam = self._lambda(s_group, [target])
            hit = self._memo[key] = torch.sum((1 - lam) * self.A_t[s_group, target] + lam * self.B_t[s_group, target])
        return hit

    # ---------------------------------------------------------------- cdc.py:240-294
    def _source_domains(self, t_group, group_idx):
        key = ("src", tuple(t_group), int(group_idx), self.initial_s_group2domain_list is None)
        hit = self._memo.get(key)
        if hit is None:
            hit = self._memo[key] = self._source_domains_uncached(t_group, group_idx)
        return list(hit)                                         # callers own (and may extend) the list

    def _source_domains_uncached(self, t_group, group_idx):
        nd = self.n_domain
        s_group = self._centers(t_group, center_num=2)
        useful = True
        P = None
        if self.initial_s_group2domain_list is not None:           # constant over the growth loop
            P = (1 - 2 * self._lambda(self.initial_s_group2domain_list[group_idx])) * torch.pow(self.w_t, 0.5)
        # constant over the growth loop (they depend on the target group only): the normalised weights and the two affinity blocks
        wt = self.w_t[t_group]
        sw = wt.sum()
        if sw != 0:
            wt = wt / sw
        A_blk, B_blk = self.A_t[:nd, t_group], self.B_t[:nd, t_group]
        while useful and len(s_group) < nd:
            lam = torch.zeros(nd, len(t_group), dtype=torch.float32)
            cands = [d for d in range(nd) if d not in s_group]
            lam[cands] = self._lambda_candidates(s_group, cands, t_group)
            J = (((1 - lam) * A_blk + lam * B_blk) * wt).sum(dim=1)
            if P is None:
                result = J
            else:
                result = J + self.p_weight * P if self.max_better else J - self.p_weight * P
            result[s_group] = float(self.default_metric_value)
            best_value, best = torch.max(result, 0) if self.max_better else torch.min(result, 0)
            useful = bool(best_value > 0) if self.max_better else bool(best_value < 0)
            if useful:
                s_group.append(int(best))
        return s_group

    # ---------------------------------------------------------------- cdc.py:121-238
    def update(self, matrix_A, matrix_B, matrix_mask, mode="iterative"):
        """matrix_A (n_domain+1, n_domain), matrix_B (n_domain+n_cluster, n_domain), matrix_mask (n_mask, n_domain): float32.
        Returns dict(A, B, mask, causal) - the transformed matrices the reference leaves on the module - and updates the lists."""
        nd, nc = self.n_domain, self.n_cluster
        self._memo = {}
        self.call_update_group += 1
        self._update_p_weight()
        A, B, M = (np.array(t, dtype=F32, copy=True) for t in (matrix_A, matrix_B, matrix_mask))
        if self.old_matrix_weight > 0 and self.old_A is not None:
            ow = F32(self.old_matrix_weight)
            A = self.old_A * ow + A * (F32(1) - ow)
            B = self.old_B * ow + B * (F32(1) - ow)
        self.old_A, self.old_B, self.old_mask = A.copy(), B.copy(), M.copy()
        d2g = np.asarray(self.domain2group_list, dtype=np.int64)
        if self.affinity_func == "minus":
            A[:-1] -= A[-1]
            B[:nd] = B[d2g + nd] - B[:nd]
            M = M - A[-1]
        elif self.affinity_func == "divide":
            A[:-1] = F32(1) - A[:-1] / A[-1]
            B[:nd] = F32(1) - B[d2g + nd] / B[:nd]
            M = F32(1) - M / A[-1]
        else:
            raise ValueError("Unknown affinity_func: " + str(self.affinity_func))
        self.A, self.B = A, B
        self.causal = np.arccos(calc_causal_matrix(M.T)).astype(F32)
        self.A_t, self.B_t, self.causal_t = torch.from_numpy(A), torch.from_numpy(B), torch.from_numpy(self.causal)

        if max(self.domain2group_list) == 0:
            # first call: k-means on the rows of the causal distance matrix (unseeded upstream: NumPy's global RNG)
            from sklearn.cluster import KMeans
            labels = KMeans(n_clusters=nc).fit(self.causal).labels_
            t_groups = [[] for _ in range(nc)]
            for i, gidx in enumerate(labels):
                t_groups[int(gidx)].append(i)
            self.domain2group_list = [int(v) for v in labels]
            self.t_group2domain_list = t_groups
            self.s_group2domain_list = [self._source_domains(t_groups[c], c) for c in range(nc)]
            self.initial_s_group2domain_list = copy.deepcopy(self.s_group2domain_list)
        else:
            t_old = self.t_group2domain_list
            queue = list(range(nd))
            t_group, s_group = [[] for _ in range(nc)], [[] for _ in range(nc)]
            metric = torch.zeros(nd, nc, dtype=torch.float32)       # upstream: torch.empty, every entry is written before it is read
            centers = [self._centers(t_old[c])[0] for c in range(nc)]
            for c in range(nc):
                t_group[c].append(centers[c])
                queue.remove(centers[c])
                metric[centers[c], :] = float(self.default_metric_value)
            pick = (lambda v: int(torch.argmax(v))) if self.max_better else (lambda v: int(torch.argmin(v)))
            if mode == "iterative":
                progressed = True
                while queue and progressed:
                    progressed = False
                    for c in range(nc):
                        s_group[c] = self._source_domains(t_group[c], c)
                    for d in queue:
                        for c in range(nc):
                            metric[d, c] = self._metric_in_source_group(d, s_group[c])
                    best = [int(v) for v in (torch.argmax(metric, dim=0) if self.max_better else torch.argmin(metric, dim=0))]
                    for c in range(nc):
                        if int(pick(metric[best[c], :])) == c:
                            progressed = True
                            t_group[c].append(best[c])
                            queue.remove(best[c])
                            metric[best[c], :] = float(self.default_metric_value)
                if queue:
                    raise ValueError("target domain_queue is not empty")
            elif mode == "greedy":
                for c in range(nc):
                    s_group[c] = self._source_domains(t_group[c], c)
                for d in queue:
                    for c in range(nc):
                        metric[d, c] = self._metric_in_source_group(d, s_group[c])
                for d in queue:
                    t_group[int(pick(metric[d, :]))].append(d)
            else:
                raise ValueError(f"unknown update_group mode {mode!r}")
            self.t_group2domain_list = t_group
            d2g_new = np.zeros(nd, dtype=np.int64)
            for c in range(nc):
                self.s_group2domain_list[c] = self._source_domains(t_group[c], c)
                d2g_new[t_group[c]] = c
            self.domain2group_list = d2g_new.tolist()
        return dict(A=A, B=B, mask=M, causal=self.causal)
