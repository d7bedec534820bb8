"""ctypes binding of libcdcmdr.so (include/cdcmdr.h).  This is the ONLY compute backend of the package: if the
shared library is missing the import of any model fails loudly - there is no CPU or PyTorch fallback.

Every wrapper takes raw device addresses (ints) and sizes exactly as the C-ABI does; `check()` turns a non-zero
status into a RuntimeError carrying cdcmdr_last_error().
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcdcmdr.so")

P = C.c_void_p
I64 = C.c_int64
I32 = C.c_int32
F32 = C.c_float
U64 = C.c_uint64
U32 = C.c_uint32
SZ = C.c_size_t
INT = C.c_int


class GemmF32(C.Structure):
    _fields_ = [("A", P), ("Bt", P), ("C", P),
                ("M", I64), ("N", I64), ("K", I64),
                ("a_rs", I64), ("a_cs", I64), ("b_rs", I64), ("b_cs", I64), ("c_rs", I64),
                ("G", I32), ("a_gs", I64), ("b_gs", I64), ("c_gs", I64),
                ("bias", P), ("bias_gs", I64),
                ("act", I32),
                ("mask", P), ("mask_rs", I64), ("mask_gs", I64), ("mask_scale", F32),
                ("drop_p", F32), ("seed_dev", P), ("salt", U32),
                ("accumulate", I32),
                ("split_k", I32), ("workspace", P)]


class GemmBf16(C.Structure):
    _fields_ = [("A", P), ("lda", I64), ("a_rows", I64), ("a_cols", I64),
                ("Bt", P), ("ldb", I64), ("b_rows", I64), ("b_cols", I64),
                ("M", I64), ("N", I64), ("K", I64),
                ("G", I32), ("a_gm", I64), ("a_gk", I64), ("b_gn", I64), ("b_gk", I64),
                ("a_mn_major", I32), ("b_mn_major", I32),
                ("bias", P), ("bias_gs", I64),
                ("n_main", I64),
                ("out_main", P), ("ld_main", I64), ("main_gn", I64),
                ("out_aux", P), ("ld_aux", I64), ("aux_gn", I64),
                ("act", I32),
                ("mask", P), ("ld_mask", I64), ("mask_gn", I64), ("mask_scale", F32),
                ("drop_p", F32), ("seed_dev", P), ("salt", U32),
                ("accumulate", I32),
                ("split_k", I32), ("aux_split_stride", I64),
                ("block_n", I32),
                ("cross_x0", P), ("cross_x", P), ("ld_cross", I64)]


class PleChain(C.Structure):
    _fields_ = [("X", P), ("ldx", I64), ("B", I64), ("K0", I32),
                ("W0", P), ("b0", P), ("W1", P), ("b1", P),
                ("nE", I32), ("d0", I32), ("d1", I32), ("n_g", I32),
                ("A0", P), ("lda0", I64), ("H", P), ("ldh", I64), ("Lg", P), ("ldg", I64),
                ("drop_p", F32), ("seed_dev", P), ("salt0", U32), ("salt1", U32)]


class MixDesc(C.Structure):
    _fields_ = [("n_gates", I32), ("n_experts", I32), ("h", I32), ("max_sel", I32),
                ("gate_col", P), ("gate_n", P), ("gate_sel", P), ("n_pairs", I32)]


class BnDesc(C.Structure):
    _fields_ = [("gamma", P), ("beta", P), ("gamma2", P), ("beta2", P),
                ("running_mean", P), ("running_var", P),
                ("save_mean", P), ("save_invstd", P),
                ("train", I32), ("relu", I32),
                ("drop_p", F32), ("seed_dev", P), ("salt", U32)]


STEP_STATE_BYTES = 48
STEP_STATE_SEED_OFFSET = 8

# name -> (restype, argtypes).  Must list every symbol include/cdcmdr.h declares (tests check this).
SIGNATURES = {
    "cdcmdr_last_error": (C.c_char_p, []),
    "cdcmdr_version": (INT, []),
    "cdcmdr_launch_count": (I64, []),
    "cdcmdr_launch_count_reset": (None, []),
    "cdcmdr_step_state_init": (INT, [P, I64, P]),
    "cdcmdr_step_tick": (INT, [P, F32, F32, F32, F32, F32, U64, P]),
    "cdcmdr_embed_gather_fwd": (INT, [P, P, P, P, P, I64, I64, INT, INT, I64, P, P]),
    "cdcmdr_embed_plan_bytes": (SZ, [I64, I64, INT]),
    "cdcmdr_embed_plan_build": (INT, [P, P, I64, INT, I64, INT, P, SZ, P]),
    "cdcmdr_embed_bwd_dense": (INT, [P, I64, P, INT, I64, INT, INT, I64, P, P]),
    "cdcmdr_embed_bwd_adam_dense_exact": (INT, [P, I64, P, INT, I64, INT, INT, I64, P, P, P, F32, P, P, P]),
    "cdcmdr_embed_bwd_adam_dense_exact_g16": (INT, [P, I64, P, INT, I64, INT, INT, I64, P, P, P, F32, P, P, P]),
    "cdcmdr_embed_bwd_adam_sparse_lazy": (INT, [P, I64, P, INT, I64, INT, INT, I64, P, P, P, F32, P, P]),
    "cdcmdr_embed_bwd_adam_sparse_lazy_reg": (INT, [P, I64, P, INT, I64, INT, INT, I64, P, P, P, F32, P, P, P, P]),
    "cdcmdr_embed_gather_peer": (INT, [P, P, P, I64, P, P, I64, I64, INT, INT, I64, P, P]),
    "cdcmdr_gemm_f32": (INT, [C.POINTER(GemmF32), P]),
    "cdcmdr_gemm_bf16_tc": (INT, [C.POINTER(GemmBf16), P]),
    "cdcmdr_gemm_bf16_tc_splits": (INT, [I64, I32]),
    "cdcmdr_gemm_bf16_tc_mode": (INT, [INT]),
    "cdcmdr_gemm_bf16_tc_profile": (INT, [P]),
    "cdcmdr_splitk_reduce": (INT, [P, I64, I32, P, I64, I64, I64, I64, I32, P]),
    "cdcmdr_transpose_bf16": (INT, [P, I64, P, I64, I64, I64, P]),
    "cdcmdr_gate_mix_fwd": (INT, [C.POINTER(MixDesc), P, I64, P, I64, P, I64, P, I64, INT, P]),
    "cdcmdr_gate_mix_bwd": (INT, [C.POINTER(MixDesc), P, I64, P, P, I64, P, I64, F32, P, I64, I64, INT, P]),
    "cdcmdr_bn_scratch_bytes": (SZ, [I64]),
    "cdcmdr_bn_fwd": (INT, [C.POINTER(BnDesc), P, I64, P, I64, INT, I64, I64, P, P]),
    "cdcmdr_bn_bwd": (INT, [C.POINTER(BnDesc), P, I64, P, I64, INT, P, I64, INT, P, I64, INT, P, P, INT, I64, I64, P, P]),
    "cdcmdr_bn_fwd_stats": (INT, [P, I64, I64, I64, P, P, P]),
    "cdcmdr_bn_fwd_apply": (INT, [C.POINTER(BnDesc), P, I64, P, I64, INT, I64, I64, I64, P, P]),
    "cdcmdr_bn_bwd_stats": (INT, [C.POINTER(BnDesc), P, I64, P, I64, INT, P, I64, INT, P, P, INT, I64, I64, P, P, P]),
    "cdcmdr_bn_bwd_apply": (INT, [C.POINTER(BnDesc), P, I64, P, I64, INT, P, I64, INT, P, I64, INT, I64, I64, I64, P, P, P]),
    "cdcmdr_rowdot_fwd": (INT, [P, I64, INT, P, P, P, I64, I64, INT, INT, P]),
    "cdcmdr_rowdot_bwd": (INT, [P, I64, INT, P, P, I64, P, I64, P, P, I64, INT, INT, P, P]),
    "cdcmdr_sigmoid_select_bce": (INT, [P, P, I64, I64, I32, I32, P, I32, P, INT, P, P, P, P, P, I64, F32, P, P]),
    "cdcmdr_sigmoid_bwd": (INT, [P, P, P, P, I64, I64, I32, P]),
    "cdcmdr_reg_l2_sum": (INT, [P, P, F32, I64, P, P, P]),
    "cdcmdr_reduce_scratch_bytes": (SZ, []),
    "cdcmdr_reg_l2_grad": (INT, [P, P, F32, F32, P, INT, I64, P]),
    "cdcmdr_relu_mask_f32": (INT, [P, I64, P, I64, P, I64, I64, I64, F32, P]),
    "cdcmdr_relu_mask": (INT, [P, I64, INT, P, I64, INT, P, I64, INT, I64, I64, F32, P]),
    "cdcmdr_adam_dense": (INT, [P, P, P, P, P, P, I64, P, P]),
    "cdcmdr_colsum": (INT, [P, I64, INT, I64, I64, P, INT, P, P]),
    "cdcmdr_colsum_scratch_bytes": (SZ, [I64]),
    "cdcmdr_cast_f32_bf16": (INT, [P, I64, P, I64, I64, I64, P]),
    "cdcmdr_cast_bf16_f32": (INT, [P, I64, P, I64, I64, I64, INT, P]),
    "cdcmdr_ewise_f32": (INT, [P, P, P, I64, INT, P]),
    "cdcmdr_ewise_group_f32": (INT, [P, P, P, I64, INT, INT, P]),
    "cdcmdr_add2d_f32": (INT, [P, I64, P, I64, I64, I64, INT, P]),
    "cdcmdr_cross_fuse_fwd": (INT, [P, P, P, INT, P, P, I64, I64, P]),
    "cdcmdr_cross_fuse_bwd": (INT, [P, P, INT, P, P, P, I64, I64, P]),
    "cdcmdr_crossmix_combine_fwd": (INT, [P, P, P, P, P, P, I64, I64, INT, P]),
    "cdcmdr_crossmix_combine_bwd": (INT, [P, P, P, P, P, P, P, P, I64, I64, INT, P]),
    "cdcmdr_tanh_fwd": (INT, [P, I64, P]),
    "cdcmdr_tanh_bwd": (INT, [P, P, I64, P]),
    "cdcmdr_softmax_rows_fwd": (INT, [P, I64, P, I64, I64, INT, P]),
    "cdcmdr_softmax_rows_bwd": (INT, [P, I64, P, I64, P, I64, I64, INT, P]),
    "cdcmdr_route_scratch_bytes": (SZ, [I64, INT]),
    "cdcmdr_route_partition": (INT, [P, I64, INT, P, P, P, P, P]),
    "cdcmdr_permute_rows": (INT, [P, I64, P, I64, I64, INT, P, I64, INT, P]),
    "cdcmdr_copy2d_batched": (INT, [P, I64, I64, P, I64, I64, I64, I64, I64, INT, P]),
    "cdcmdr_bce_segments_scratch_bytes": (SZ, [INT]),
    "cdcmdr_bce_segments": (INT, [P, I64, P, INT, P, INT, P, P, P]),
    "cdcmdr_domain_to_group": (INT, [P, I64, INT, INT, P, INT, P, P]),
    "cdcmdr_attn_fwd": (INT, [P, I64, P, I64, P, I64, INT, INT, INT, F32, F32, P, U32, P]),
    "cdcmdr_attn_bwd": (INT, [P, I64, P, P, I64, P, I64, I64, INT, INT, INT, F32, F32, P, U32, P]),
    "cdcmdr_attn_pool_fwd": (INT, [P, P, P, I64, INT, I64, I64, P]),
    "cdcmdr_attn_pool_scratch_bytes": (C.c_size_t, [I64, I64]),
    "cdcmdr_attn_pool_bwd": (INT, [P, P, P, I64, P, P, I64, I64, P, P]),
    "cdcmdr_ple_chain_ok": (INT, [I32, I32, I32, I32]),
    "cdcmdr_ple_chain_fwd": (INT, [C.POINTER(PleChain), P]),
    "cdcmdr_ple_chain_profile": (INT, [P]),
    "cdcmdr_auc_logloss_scratch_bytes": (SZ, [I64, I32]),
    "cdcmdr_auc_logloss": (INT, [P, P, INT, P, INT, I64, I32, P, P, P]),
    "cdcmdr_peer_allreduce_bytes": (SZ, [INT, I64]),
    "cdcmdr_peer_allreduce_f64": (INT, [P, INT, INT, P, P, I64, I64, P, P]),
    "cdcmdr_peer_allreduce_f32_chunk": (I64, [INT, I64]),
    "cdcmdr_peer_allreduce_f32": (INT, [P, P, P, P, P, INT, INT, P, P, I64, P, P]),
    "cdcmdr_peer_barrier": (INT, [P, INT, INT, INT, INT, P, P]),
    "cdcmdr_dp_push_ids": (INT, [P, I64, INT, P, P, INT, INT, P]),
    "cdcmdr_dp_gather_push": (INT, [P, P, P, I64, P, INT, I64, INT, I64, INT, INT, INT, INT, P, P]),
    "cdcmdr_dp_push_grads": (INT, [P, I64, I64, INT, INT, P, INT, P, INT, INT, P]),
    "cdcmdr_attn_fwd_bf16": (INT, [P, I64, P, I64, I64, INT, INT, INT, F32, F32, P, U32, P]),
    "cdcmdr_attn_bwd_bf16": (INT, [P, I64, P, I64, P, I64, I64, INT, INT, INT, F32, F32, P, U32, P]),
    "cdcmdr_attn_pool_fwd_bf16": (INT, [P, P, P, I64, INT, I64, I64, P]),
    "cdcmdr_attn_pool_bwd_bf16": (INT, [P, P, P, I64, P, P, I64, I64, P, P]),
}

# entry points that return a status code (everything that launches work)
_STATUS = {k for k, (r, _) in SIGNATURES.items() if r is INT and k not in ("cdcmdr_version", "cdcmdr_gemm_bf16_tc_splits", "cdcmdr_gemm_bf16_tc_mode", "cdcmdr_ple_chain_ok")}


class CdcmdrError(RuntimeError):
    pass


class Lib:
    """Thin object over the loaded shared library: `lib.<name-without-prefix>(*args)` calls the C entry point and
    raises CdcmdrError on a non-zero status."""

    def __init__(self, path: str = LIB_PATH):
        if not os.path.exists(path):
            raise CdcmdrError(
                f"{path} is missing: build it with `python __graft_entry__.py build` (nvcc, sm_100a). "
                "This package has no CPU / PyTorch fallback.")
        self.path = path
        self._dll = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(self._dll, name)          # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
            short = name[len("cdcmdr_"):]
            setattr(self, short, self._wrap(name, fn) if name in _STATUS else fn)

    def _wrap(self, name, fn):
        last_error = self._dll.cdcmdr_last_error

        def call(*args):
            rc = fn(*args)
            if rc != 0:
                raise CdcmdrError(f"{name} failed ({rc}): {last_error().decode(errors='replace')}")
            return rc
        call.__name__ = name
        return call


_LIB = None


def load() -> Lib:
    """The process-wide library handle (tests may install an emulator with `install`)."""
    global _LIB
    if _LIB is None:
        _LIB = Lib()
    return _LIB


def install(lib) -> None:
    """Replace the process-wide handle.  Used ONLY by tests/ to run the host logic against the host-memory
    emulator of the C-ABI (oracle/host_abi.py); never called by the package."""
    global _LIB
    _LIB = lib
