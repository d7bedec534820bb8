"""AutoInt (reference model/autoint.py:10-64) on libcdcmdr.so.

    atten_x = atten_embedding(embed_x.view(B, F, E)) ; att_layer_num x nn.MultiheadAttention ; (+ V_res_embedding(embed_x))
    y = sigmoid( linear(e) + [relu(atten).view(B, F*A) | MLP(e)] . w )            w = dnn_linear.weight (1, F*A + h), no bias

The attention stack is the same block BaseModel.atten_forward runs (atten.AttnBlock: bf16 token matrices through the tcgen05 GEMM
and the mma.sync attention core on the tensor-core path, fp32 kernels otherwise) with its ReLU + Linear head reading the first
F*A columns of `dnn_linear`; the MLP branch is an `MlpGroup`, its share of the final Linear a row-dot with the remaining columns.
Single-logit model: forward returns (B,) probabilities; the fused step is `train_step(x, y, optimizer, mode="col", col=0)`."""
from __future__ import annotations

from torch import nn

from .core import Mat
from .dcn import _CrossModel
from .layer import MultiLayerPerceptron, mlp_group_names, precision_of
from .runtime import MlpGroup


class AutoInt(_CrossModel):
    def __init__(self, feature_dims, embed_dim, atten_embed_dim=None, att_layer_num=3, att_head_num=2, att_res=True,
                 mlp_dims=(256, 128), dropout=0.2, l2_reg_embedding=1e-5, l2_reg_linear=1e-5, l2_reg_dnn=1e-5, config=None):
        super().__init__(feature_dims, embed_dim, l2_reg_embedding=l2_reg_embedding, l2_reg_linear=l2_reg_linear)
        self.model_name = 'autoint'
        if len(mlp_dims) <= 0 and att_layer_num <= 0:
            raise ValueError("Either MLP hidden_layer or att_layer_num must > 0")
        if len(mlp_dims) <= 0 or att_layer_num <= 0:
            raise NotImplementedError("cdcmdr: AutoInt needs at least one attention layer and one MLP layer")
        if atten_embed_dim is None:
            atten_embed_dim = embed_dim
        if self.field_num > 32:
            raise NotImplementedError("cdcmdr: the attention core handles at most 32 field tokens")
        # module creation order = the reference's (autoint.py:33-47): parameter initialisation consumes torch's RNG in this order
        self.atten_embedding = nn.Linear(embed_dim, atten_embed_dim)
        self.atten_output_dim = self.embedding.output_dim0 * atten_embed_dim
        self.att_res = att_res
        self.mlp_dims = tuple(mlp_dims)
        self.dnn = MultiLayerPerceptron(self.embed_output_dim, mlp_dims, dropout, output_layer=False)
        self.self_attns = nn.ModuleList([nn.MultiheadAttention(atten_embed_dim, att_head_num, dropout=dropout)
                                         for _ in range(att_layer_num)])
        if self.att_res:
            self.V_res_embedding = nn.Linear(embed_dim, atten_embed_dim)
        self.dnn_linear = nn.Linear(mlp_dims[-1] + self.atten_output_dim, 1, bias=False)
        self.output_layer = nn.Sigmoid()
        self.add_regularization_weight(self.reg_filter("dnn"), l2=l2_reg_dnn)
        self._att_geom = (atten_embed_dim, int(att_head_num), int(att_layer_num))
        self._att_head = ("dnn_linear.weight", 0)
        self.use_atten = True
        self._mlp_names, blk, bufs = mlp_group_names(["dnn"], self.dnn, "dnn")
        self._finalize(blk, bufs, precision=precision_of(config), dropout=dropout)

    def _on_runtime_built(self):
        self._mlp = MlpGroup(self._rt, "dnn", 1, self.embed_output_dim, self.mlp_dims, self._mlp_names, bn=True, out_layer=False,
                             in_groups=None)

    def _program_fwd(self, ws, X: Mat, B, train):
        rt, FA, h = self._rt, self.atten_output_dim, self.mlp_dims[-1]
        lin = self._lin_fwd(ws, X, B)                                   # FeaturesLinear (layer.py:115-126)
        self._att.fwd(ws, self._att_x(ws, X, B), B, lin, train)         # += relu(cross_term) . dnn_linear.weight[:, :F*A]
        mlp_out = self._mlp.fwd(ws, X, B, train)
        logit = ws.mat("head.logit", B, 1)
        rt.ops.rowdot_fwd(mlp_out, rt.w("dnn_linear.weight", FA), None, logit, B, 1, h)
        return logit, lin

    def _program_bwd(self, ws, X: Mat, B, train, dlogits: Mat):
        rt, D, FA, h = self._rt, self.embed_output_dim, self.atten_output_dim, self.mlp_dims[-1]
        mlp_out = self._mlp._act(ws, len(self.mlp_dims) - 1, B)
        dmlp = ws.mat("head.dmlp", B, h)
        rt.ops.rowdot_bwd(mlp_out, rt.w("dnn_linear.weight", FA), dlogits, dmlp, rt.g("dnn_linear.weight", FA), None, B, 1, h)
        dX = ws.mat("dX", B, D)
        self._mlp.bwd(ws, X, dmlp, B, train, dX, post_act_grad=True)
        self._att.bwd(ws, self._att_x(ws, X, B), B, self._dlin_mat(ws, B), dX, train)
        self._lin_bwd(ws, X, B, dX)
        return dX
