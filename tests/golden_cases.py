"""Shared description of the golden fixtures (tests/golden/make_golden.py) for oracle and GPU tests."""
import os
import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIELD_DIMS = np.array([7, 5, 11, 4, 9, 6], dtype=np.int64)
E = 4
L2 = dict(l2_reg_embedding=1e-3, l2_reg_linear=1e-3, l2_reg_dnn=1e-3, l2_reg_cross=1e-3)
T = 3

# name -> (model kind, ctor kwargs, selection mode, steps)
CASES = {
    "ple": ("ple", dict(n_tower=T, n_expert_specific=2, n_expert_shared=1, expert_dims=((16, 8), (8,)), tower_dims=(8, 4)), "gather", 3),
    "mmoe": ("mmoe", dict(n_tower=T, n_expert=3, expert_dims=(16, 8), tower_dims=(8, 4)), "gather", 3),
    "dcn": ("dcn", dict(n_cross_layers=3, mlp_dims=(16, 8)), "single", 3),
    "dcnv2_mix_parallel": ("dcnv2", dict(n_cross_layers=2, mlp_dims=(16, 8), low_rank=4, num_experts=3), "single", 3),
    "dcnv2_mix_stacked": ("dcnv2", dict(n_cross_layers=2, mlp_dims=(16, 8), model_structure="stacked", low_rank=4, num_experts=3), "single", 2),
    "star": ("star", dict(n_tower=T, tower_dims=(16, 8)), "gather", 3),
    "star_grouped": ("star", dict(n_tower=T, tower_dims=(16, 8)), "star_grouped", 3),
}
for _base, _kw in (("ple", CASES["ple"][1]), ("mmoe", CASES["mmoe"][1]), ("star", CASES["star"][1])):
    for _mode, _sel in (("warmup", "warmup"), ("split_domain", "split_domain"), ("split_gather", "split_gather")):
        CASES[f"cdc_{_base}_{_mode}"] = (_base, _kw, _sel, 2)

# the field self-attention block (config.use_atten; tests/golden/make_golden_atten.py): the same models with the
# block switched on; the numpy model oracle is pinned on them too (Base.enable_atten, tests/test_oracle_golden.py)
ATTEN = {"ple_atten": dict(atten_embed_dim=8, att_layer_num=2, att_head_num=2, att_res=True),
         "mmoe_atten": dict(atten_embed_dim=8, att_layer_num=3, att_head_num=2, att_res=False),
         "star_atten": dict(atten_embed_dim=8, att_layer_num=2, att_head_num=2, att_res=True),
         "star_grouped_atten": dict(atten_embed_dim=8, att_layer_num=2, att_head_num=2, att_res=True)}
ATTEN_CASES = {"ple_atten": CASES["ple"], "mmoe_atten": CASES["mmoe"], "star_atten": CASES["star"],
               "star_grouped_atten": CASES["star_grouped"]}
# AutoInt (model/autoint.py; tests/golden/make_golden_autoint.py): its L2 set has no l2_reg_cross
AUTOINT_CASES = {"autoint": ("autoint", dict(atten_embed_dim=8, att_layer_num=3, att_head_num=2, att_res=True, mlp_dims=(16, 8)), "single", 3),
                 "autoint_nores": ("autoint", dict(atten_embed_dim=None, att_layer_num=2, att_head_num=1, att_res=False, mlp_dims=(16,)), "single", 3)}
ALL_CASES = {**CASES, **ATTEN_CASES, **AUTOINT_CASES}


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def state(gold, step, strip="base_model_instance."):
    pre = f"sd{step}."
    out = {}
    for k, v in gold.items():
        if k.startswith(pre):
            kk = k[len(pre):]
            out[kk[len(strip):] if kk.startswith(strip) else kk] = v
    return out


def strip(k, prefix="base_model_instance."):
    return k[len(prefix):] if k.startswith(prefix) else k
