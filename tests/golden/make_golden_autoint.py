#!/usr/bin/env python
"""Golden fixtures for AutoInt (reference model/autoint.py:10-64), produced by the UNMODIFIED reference:

    python tests/golden/make_golden_autoint.py

Two geometries (attention width 8 with 2 heads and the V_res residual, 3 layers; attention width = embed_dim with 1 head, no
residual, 2 layers), dropout 0, three steps of the reference's loop body - same layout as make_golden.py (run_case)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (installs the import shim, imports the reference models)
from model.autoint import AutoInt  # noqa: E402

L2 = dict(l2_reg_embedding=1e-3, l2_reg_linear=1e-3, l2_reg_dnn=1e-3)
CASES = {"autoint": dict(atten_embed_dim=8, att_layer_num=3, att_head_num=2, att_res=True, mlp_dims=(16, 8)),
         "autoint_nores": dict(atten_embed_dim=None, att_layer_num=2, att_head_num=1, att_res=False, mlp_dims=(16,))}


def main():
    torch.manual_seed(2005)
    torch.set_num_threads(1)
    rng = np.random.default_rng(2005)
    batches = [G.make_batch(rng, 24, 3) for _ in range(2)]
    for name, kw in CASES.items():
        m = AutoInt(G.FIELD_DIMS, G.E, dropout=0.0, **kw, **L2)
        G.run_case(name, m, G.fwd_single, batches, 3)


if __name__ == "__main__":
    main()
