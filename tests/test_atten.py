"""The field self-attention block (config.use_atten - ON in the reference's stock config.py:24-28; model/layer.py:58-84; SURVEY
§8f N3) against fixtures produced by the unmodified reference with the block switched on (tests/golden/make_golden_atten.py):
PLE with 2 attention layers + the V_res residual, MMoE with 3 layers without it; predictions, losses, every gradient, the
state_dict after each of three Adam steps, eval forward - through the fused step and through loss.backward()."""
import numpy as np
import pytest
import torch

import cdcmdr_b200 as cm
from oracle.host_abi import HostABI
from tests.golden_cases import ATTEN, AUTOINT_CASES, load, state
from tests.util import build_model, run_golden_case


@pytest.fixture
def emulator():
    old = cm._lib._LIB
    cm._lib.install(HostABI())
    yield
    cm._lib.install(old)


@pytest.mark.parametrize("path", ["fused", "autograd"])
@pytest.mark.parametrize("name", sorted(ATTEN))
def test_attention_block_host_logic(name, path, emulator):
    run_golden_case(name, "cpu", path=path)


@pytest.mark.parametrize("name", sorted(ATTEN))
def test_attention_state_dict_keys_match_reference(name, emulator):
    m = build_model(name)
    ref = state(load(name), 0)
    mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert set(mine) == set(ref)
    for k, v in ref.items():
        assert mine[k] == tuple(v.shape), k
    assert any(k.startswith("self_attns.0.in_proj_weight") for k in mine) and "atten_linear.weight" in mine


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["fused", "autograd"])
@pytest.mark.parametrize("name", sorted(ATTEN))
def test_attention_block_gpu(name, path):
    run_golden_case(name, "cuda", path=path)


# ---- AutoInt (model/autoint.py:10-64): the same attention stack + an MLP branch under one bias-free Linear; fixtures from the
# unmodified reference (tests/golden/make_golden_autoint.py)
@pytest.mark.parametrize("path", ["fused", "autograd"])
@pytest.mark.parametrize("name", sorted(AUTOINT_CASES))
def test_autoint_host_logic(name, path, emulator):
    run_golden_case(name, "cpu", path=path)


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["fused", "autograd"])
@pytest.mark.parametrize("name", sorted(AUTOINT_CASES))
def test_autoint_gpu(name, path):
    run_golden_case(name, "cuda", path=path)


@pytest.mark.gpu
def test_autoint_bf16_path_gpu():
    """AutoInt on the tensor-core path (bf16 token matrices, tcgen05 projections, mma.sync attention core) against its fp32 path on
    identical weights: eval logits 2e-2, first step's BCE 2e-2, attention and head weights train."""
    from tests.golden_cases import FIELD_DIMS
    res, sd = {}, None
    rng = np.random.default_rng(6)
    B = 300
    x = torch.from_numpy(np.stack([rng.integers(0, d, size=B) for d in FIELD_DIMS], axis=1).astype(np.int32)).cuda()
    y = torch.from_numpy((rng.random(B) < 0.3).astype(np.int16)).cuda()
    for prec in ("fp32", "bf16"):
        class Cfg:
            cdcmdr_precision = prec
        torch.manual_seed(12)
        m = cm.AutoInt(FIELD_DIMS, 8, atten_embed_dim=32, att_layer_num=2, att_head_num=2, att_res=True, mlp_dims=(32, 16), dropout=0.0,
                       l2_reg_embedding=1e-3, l2_reg_linear=1e-3, l2_reg_dnn=1e-3, config=Cfg())
        if sd is None:
            sd = {k: v.clone() for k, v in m.state_dict().items()}
        else:
            m.load_state_dict(sd, strict=True)
        m = m.cuda().eval()
        with torch.no_grad():
            p = m(x).cpu().numpy().astype(np.float64)
        m.train()
        opt = cm.Adam(m.parameters(), lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)
        w0 = m.state_dict()["dnn_linear.weight"].clone()
        bce = m.step_losses(m.train_step(x, y, opt, mode="col", col=0))[1]
        dw = (m.state_dict()["dnn_linear.weight"] - w0).abs().reshape(-1)
        moved = float(min(dw[:6 * 32].mean(), dw[6 * 32:].mean()))     # attention part and MLP part of the head (a dead ReLU column may rest)
        res[prec] = (np.log(p) - np.log1p(-p), bce, moved)
    l32, l16 = res["fp32"][0], res["bf16"][0]
    assert np.abs(l16 - l32).max() <= 2e-2 * max(1.0, float(np.abs(l32).max()))
    assert abs(res["bf16"][1] - res["fp32"][1]) <= 2e-2 * abs(res["fp32"][1])
    assert res["fp32"][2] > 1e-4 and res["bf16"][2] > 1e-4          # both parts of the head trained


def _bf16_pair(kind, device):
    """the same weights on the fp32 path and on the bf16 tensor-core path (layer widths multiples of 8, as that path needs)"""
    from tests.golden_cases import FIELD_DIMS, L2

    class Cfg:
        use_atten = True; use_dcn = False; atten_embed_dim = 16; att_layer_num = 2; att_head_num = 2; att_res = True
        ple_n_expert_specific = 2; ple_n_expert_shared = 1; mmoe_n_expert = 3

    def build(prec):
        cfg = Cfg(); cfg.cdcmdr_precision = prec
        torch.manual_seed(11)
        if kind == "ple":
            return cm.PLE(FIELD_DIMS, 8, 3, 2, 1, ((32, 16), (16,)), (16, 8), dropout=0.0, config=cfg, **L2)
        return cm.MMoE(FIELD_DIMS, 8, 3, 3, (32, 16), (16, 8), dropout=0.0, config=cfg, **L2)
    m32 = build("fp32")
    m16 = build("bf16")
    m16.load_state_dict(m32.state_dict(), strict=True)
    rng = np.random.default_rng(4)
    B = 200
    x = torch.from_numpy(np.stack([rng.integers(0, d, size=B) for d in FIELD_DIMS], axis=1).astype(np.int32)).to(device)
    y = torch.from_numpy((rng.random(B) < 0.3).astype(np.int16)).to(device)
    g = torch.from_numpy(rng.integers(0, 3, size=B).astype(np.int64)).to(device)
    return m32.to(device), m16.to(device), x, y, g


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["ple", "mmoe"])
def test_attention_block_gpu_bf16_path(kind):
    """bf16 tensor-core path for the experts / gates / towers with the attention block in fp32 on the fp32 copy of the embeddings
    (one gather writes both): predictions and the first step's loss against the fp32 path on the same weights, 2e-2 on logits"""
    m32, m16, x, y, g = _bf16_pair(kind, "cuda")
    res = {}
    for prec, m in (("fp32", m32), ("bf16", m16)):
        w0 = m.state_dict()["atten_linear.weight"].cpu().numpy().copy()
        m.eval()
        with torch.no_grad():
            pred = m(x).cpu().numpy()
        m.train()
        opt = cm.Adam(m.parameters(), lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-8)
        out = m.train_step(x, y, opt, mode="gather", sel=g)
        _, bce, _ = m.step_losses(out)
        w1 = m.state_dict()["atten_linear.weight"].cpu().numpy()
        res[prec] = (pred, bce, np.abs(w1 - w0).max())
    logit = lambda p: np.log(p) - np.log1p(-p)   # noqa: E731
    l32, l16 = logit(res["fp32"][0].astype(np.float64)), logit(res["bf16"][0].astype(np.float64))
    assert np.abs(l16 - l32).max() <= 2e-2 * max(1.0, float(np.abs(l32).max()))
    assert abs(res["bf16"][1] - res["fp32"][1]) <= 2e-2 * abs(res["fp32"][1])
    assert res["fp32"][2] > 1e-4 and res["bf16"][2] > 1e-4                       # the block's weights did train on both paths


def _attn_case(B, L, H, dh, seed):
    rng = np.random.default_rng(seed)
    A = H * dh
    qkv = rng.standard_normal((B * L, 3 * A)).astype(np.float32)
    dout = rng.standard_normal((B * L, A)).astype(np.float32)
    z = rng.standard_normal((B, L * A)).astype(np.float32)
    w = rng.standard_normal(L * A).astype(np.float32)
    dlin = rng.standard_normal((B, 3)).astype(np.float32)
    lin0 = rng.standard_normal((B, 3)).astype(np.float32)
    return qkv, dout, z, w, dlin, lin0


@pytest.mark.gpu
@pytest.mark.parametrize("B,L,H,dh", [(5, 6, 2, 4), (300, 23, 2, 32), (64, 32, 4, 16), (1000, 26, 1, 64), (3, 1, 2, 8), (130, 16, 2, 40)])
def test_attention_entry_points_gpu(B, L, H, dh):
    """cdcmdr_attn_fwd / _bwd / _pool_fwd / _pool_bwd on the GPU against the numpy restatement (and, for the core, torch's own
    nn.functional.scaled_dot_product_attention math in float64)"""
    lib, emu = cm._lib.load(), HostABI()
    A = H * dh
    qkv, dout, z, w, dlin, lin0 = _attn_case(B, L, H, dh, B + L)
    scale = 1.0 / np.sqrt(dh)
    # --- emulator (host)
    h_out, h_p, h_dqkv = np.zeros((B * L, A), np.float32), np.zeros(B * H * L * L, np.float32), np.zeros((B * L, 3 * A), np.float32)
    emu.attn_fwd(qkv.ctypes.data, 3 * A, h_out.ctypes.data, A, h_p.ctypes.data, B, L, H, dh, scale, 0.0, None, 0, 0)
    emu.attn_bwd(qkv.ctypes.data, 3 * A, h_p.ctypes.data, dout.ctypes.data, A, h_dqkv.ctypes.data, 3 * A, B, L, H, dh, scale, 0.0, None, 0, 0)
    # --- float64 cross-check of the emulator itself with torch autograd
    t = torch.from_numpy(qkv).double().requires_grad_(True)
    q, k, v = (t[:, i * A:(i + 1) * A].reshape(B, L, H, dh).transpose(1, 2) for i in range(3))
    o = torch.softmax(q @ k.transpose(-1, -2) * scale, dim=-1) @ v
    o = o.transpose(1, 2).reshape(B * L, A)
    o.backward(torch.from_numpy(dout).double())
    assert np.abs(h_out - o.detach().numpy()).max() <= 1e-4
    assert np.abs(h_dqkv - t.grad.numpy()).max() <= 1e-4
    # --- GPU
    d = lambda a: torch.from_numpy(a).cuda()   # noqa: E731
    g_qkv, g_dout = d(qkv), d(dout)
    g_out, g_p, g_dqkv = torch.zeros(B * L, A, device="cuda"), torch.zeros(B * H * L * L, device="cuda"), torch.zeros(B * L, 3 * A, device="cuda")
    lib.attn_fwd(g_qkv.data_ptr(), 3 * A, g_out.data_ptr(), A, g_p.data_ptr(), B, L, H, dh, scale, 0.0, None, 0, 0)
    lib.attn_bwd(g_qkv.data_ptr(), 3 * A, g_p.data_ptr(), g_dout.data_ptr(), A, g_dqkv.data_ptr(), 3 * A, B, L, H, dh, scale, 0.0, None, 0, 0)
    torch.cuda.synchronize()
    assert np.abs(g_out.cpu().numpy() - h_out).max() <= 1e-4
    assert np.abs(g_p.cpu().numpy() - h_p).max() <= 5e-6
    assert np.abs(g_dqkv.cpu().numpy() - h_dqkv).max() <= 1e-4
    # --- head
    n = L * A
    h_lin, h_dz, h_dw = lin0.copy(), np.zeros((B, n), np.float32), np.zeros(n, np.float32)
    emu.attn_pool_fwd(z.ctypes.data, w.ctypes.data, h_lin.ctypes.data + 4, 3, 1, B, n, 0)
    emu.attn_pool_bwd(z.ctypes.data, w.ctypes.data, dlin.ctypes.data + 8, 3, h_dz.ctypes.data, h_dw.ctypes.data, B, n, None, 0)
    g_z, g_w, g_lin, g_dlin = d(z), d(w), d(lin0.copy()), d(dlin)
    g_dz, g_dw = torch.zeros(B, n, device="cuda"), torch.zeros(n, device="cuda")
    sc = torch.empty(lib.attn_pool_scratch_bytes(B, n), dtype=torch.uint8, device="cuda")
    lib.attn_pool_fwd(g_z.data_ptr(), g_w.data_ptr(), g_lin.data_ptr() + 4, 3, 1, B, n, 0)
    lib.attn_pool_bwd(g_z.data_ptr(), g_w.data_ptr(), g_dlin.data_ptr() + 8, 3, g_dz.data_ptr(), g_dw.data_ptr(), B, n, sc.data_ptr(), 0)
    torch.cuda.synchronize()
    ref_scale = max(1.0, float(np.abs(h_lin).max()))
    assert np.abs(g_lin.cpu().numpy() - h_lin).max() <= 1e-5 * ref_scale * np.sqrt(n)
    assert np.array_equal(g_lin.cpu().numpy()[:, [0, 2]], lin0[:, [0, 2]])            # neighbours of the strided column untouched
    assert np.array_equal(g_dz.cpu().numpy(), h_dz)
    assert np.abs(g_dw.cpu().numpy() - h_dw).max() <= 1e-5 * max(1.0, float(np.abs(h_dw).max()))


@pytest.mark.gpu
def test_attention_dropout_statistics_gpu():
    """attention-weight dropout (nn.MultiheadAttention(dropout=p), train mode): kept fraction, 1/(1-p) rescale, the backward uses the
    same mask (dV of a dropped weight is zero), repeatable for the same seed"""
    lib = cm._lib.load()
    B, L, H, dh, p = 400, 23, 2, 32, 0.25
    A = H * dh
    scale = 1.0 / np.sqrt(dh)
    qkv = torch.zeros(B * L, 3 * A, device="cuda")                                    # q = k = 0 -> uniform softmax 1/L
    qkv[:, 2 * A:] = 1.0                                                              # v = 1 -> out = sum of kept, rescaled weights
    seed = torch.tensor([12345], dtype=torch.int64, device="cuda")
    outs = []
    for _ in range(2):
        out, pr = torch.zeros(B * L, A, device="cuda"), torch.zeros(B * H * L * L, device="cuda")
        lib.attn_fwd(qkv.data_ptr(), 3 * A, out.data_ptr(), A, pr.data_ptr(), B, L, H, dh, scale, p, seed.data_ptr(), 7, 0)
        torch.cuda.synchronize()
        outs.append(out.cpu().numpy())
    assert np.array_equal(outs[0], outs[1])
    o = outs[0].reshape(B, L, H, dh)
    assert np.abs(o - o[..., :1]).max() == 0                                          # the mask is per (pair, i, j), not per d
    kept = o[..., 0] * (1 - p) * L                                                    # number of kept weights in the row
    assert np.abs(kept - np.round(kept)).max() < 1e-3
    frac = kept.sum() / (B * L * H * L)
    assert abs(frac - (1 - p)) < 0.01, frac
    assert np.allclose(pr.cpu().numpy(), 1.0 / L, atol=1e-6)                          # probs are saved before the dropout


@pytest.mark.gpu
@pytest.mark.parametrize("B,L,H,dh", [(5, 6, 2, 4), (300, 23, 2, 32), (64, 32, 4, 16), (1000, 26, 1, 64), (3, 1, 2, 8), (130, 16, 2, 40)])
def test_attention_bf16_entry_points_gpu(B, L, H, dh):
    """cdcmdr_attn_fwd_bf16 / _bwd_bf16 / _pool_fwd_bf16 / _pool_bwd_bf16 (the tensor-core path's token matrices are bf16; the
    backward recomputes the softmax) against the numpy restatement on identical bf16 inputs: outputs within one bf16 ulp of the
    tensor scale (fp32 arithmetic inside, summation order differs), head gradient as the fp32 kernels."""
    from oracle.host_abi import f32_to_bf16, bf16_to_f32
    lib, emu = cm._lib.load(), HostABI()
    A = H * dh
    qkv, dout, z, w, dlin, lin0 = _attn_case(B, L, H, dh, B + L + 1)
    scale = 1.0 / np.sqrt(dh)
    b16 = lambda a: f32_to_bf16(a).reshape(a.shape)                                  # noqa: E731
    qkv16, dout16, z16 = b16(qkv), b16(dout), b16(z)
    h_out, h_dqkv = np.zeros((B * L, A), np.uint16), np.zeros((B * L, 3 * A), np.uint16)
    emu.attn_fwd_bf16(qkv16.ctypes.data, 3 * A, h_out.ctypes.data, A, B, L, H, dh, scale, 0.0, None, 0, 0)
    emu.attn_bwd_bf16(qkv16.ctypes.data, 3 * A, dout16.ctypes.data, A, h_dqkv.ctypes.data, 3 * A, B, L, H, dh, scale, 0.0, None, 0, 0)
    d = lambda a: torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a).cuda()   # noqa: E731
    g_qkv, g_dout = d(qkv16), d(dout16)
    g_out = torch.zeros(B * L, A, dtype=torch.int16, device="cuda")
    g_dqkv = torch.zeros(B * L, 3 * A, dtype=torch.int16, device="cuda")
    lib.attn_fwd_bf16(g_qkv.data_ptr(), 3 * A, g_out.data_ptr(), A, B, L, H, dh, scale, 0.0, None, 0, 0)
    lib.attn_bwd_bf16(g_qkv.data_ptr(), 3 * A, g_dout.data_ptr(), A, g_dqkv.data_ptr(), 3 * A, B, L, H, dh, scale, 0.0, None, 0, 0)
    torch.cuda.synchronize()
    for got, ref, what in ((g_out, h_out, "out"), (g_dqkv, h_dqkv, "dqkv")):
        a, b = bf16_to_f32(got.cpu().numpy().view(np.uint16)), bf16_to_f32(ref)
        sc_ = max(float(np.abs(b).max()), 1e-30)
        assert float(np.abs(a - b).max()) <= 2 ** -7 * sc_, (what, float(np.abs(a - b).max()), sc_)
        assert float((np.abs(a - b) > 2 ** -8 * sc_).mean()) < 0.02, what      # P and dS are rounded to bf16 for the tensor-core products
    # --- head
    n = L * A
    h_lin, h_dz, h_dw = lin0.copy(), np.zeros((B, n), np.uint16), np.zeros(n, np.float32)
    emu.attn_pool_fwd_bf16(z16.ctypes.data, w.ctypes.data, h_lin.ctypes.data + 4, 3, 1, B, n, 0)
    emu.attn_pool_bwd_bf16(z16.ctypes.data, w.ctypes.data, dlin.ctypes.data + 8, 3, h_dz.ctypes.data, h_dw.ctypes.data, B, n, None, 0)
    g_z, g_w, g_lin, g_dlin = d(z16), d(w), d(lin0.copy()), d(dlin)
    g_dz, g_dw = torch.zeros(B, n, dtype=torch.int16, device="cuda"), torch.zeros(n, device="cuda")
    sc = torch.empty(lib.attn_pool_scratch_bytes(B, n), dtype=torch.uint8, device="cuda")
    lib.attn_pool_fwd_bf16(g_z.data_ptr(), g_w.data_ptr(), g_lin.data_ptr() + 4, 3, 1, B, n, 0)
    lib.attn_pool_bwd_bf16(g_z.data_ptr(), g_w.data_ptr(), g_dlin.data_ptr() + 8, 3, g_dz.data_ptr(), g_dw.data_ptr(), B, n, sc.data_ptr(), 0)
    torch.cuda.synchronize()
    ref_scale = max(1.0, float(np.abs(h_lin).max()))
    assert np.abs(g_lin.cpu().numpy() - h_lin).max() <= 1e-5 * ref_scale * np.sqrt(n)
    assert np.array_equal(g_dz.cpu().numpy().view(np.uint16), h_dz)
    assert np.abs(g_dw.cpu().numpy() - h_dw).max() <= 1e-5 * max(1.0, float(np.abs(h_dw).max()))
