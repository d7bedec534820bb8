#!/usr/bin/env python
"""Golden fixtures for the field self-attention block (config.use_atten, reference model/layer.py:58-84), produced
by the UNMODIFIED reference:   python tests/golden/make_golden_atten.py

PLE, MMoE and STAR with use_atten=True (atten_embed_dim 8, 2 heads; PLE / STAR: 2 attention layers with the V_res residual, MMoE:
3 layers without it; STAR in its all-rows and row-routed modes), dropout 0, three steps of the reference's loop body - same layout as make_golden.py (run_case): inputs, initial
state_dict, per-step predictions / losses, first-step gradients, state_dict after every step, eval-mode forward."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (installs the import shim, imports the reference models)


def cfg(layers, res):
    c = G.Cfg()
    c.use_atten, c.atten_embed_dim, c.att_layer_num, c.att_head_num, c.att_res = True, 8, layers, 2, res
    return c


def main():
    torch.manual_seed(2003)
    torch.set_num_threads(1)
    rng = np.random.default_rng(2003)
    B, T = 24, 3
    batches = [G.make_batch(rng, B, T) for _ in range(2)]
    m = G.PLE(G.FIELD_DIMS, G.E, T, 2, 1, ((16, 8), (8,)), (8, 4), dropout=0.0, config=cfg(2, True), **G.L2)
    G.run_case("ple_atten", m, G.fwd_multi, batches, 3)
    m = G.MMoE(G.FIELD_DIMS, G.E, T, 3, (16, 8), (8, 4), dropout=0.0, config=cfg(3, False), **G.L2)
    G.run_case("mmoe_atten", m, G.fwd_multi, batches, 3)
    # STAR: every tower on every row, and the row-routed mode (the block's output is indexed by the tower's mask, star.py:103-107)
    for name, fwd in (("star_atten", G.fwd_multi), ("star_grouped_atten", G.fwd_star_grouped)):
        m = G.STAR(G.FIELD_DIMS, G.E, T, (16, 8), domain_idx=3, dropout=0.0, config=cfg(2, True), device="cpu", **G.L2)
        with torch.no_grad():
            m.shared_bn_weight.uniform_(0.5, 1.5); m.shared_bn_bias.normal_(0, 0.1)
            for dn in m.domain_norm:
                dn.weight.uniform_(0.5, 1.5); dn.bias.normal_(0, 0.1)
        G.run_case(name, m, fwd, batches, 3)


if __name__ == "__main__":
    main()
