"""The C-ABI shared library loads and exports exactly the entry points include/cdcmdr.h declares, and the ctypes
binding lists the same set (no compute calls: runs without a GPU)."""
import os
import re
import subprocess

import cdcmdr_b200 as cm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared():
    src = open(os.path.join(ROOT, "include", "cdcmdr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(cdcmdr_[a-z0-9_]+)\s*\(", src))


def ensure_built():
    if not os.path.exists(cm._lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return cm._lib.LIB_PATH


def test_header_binding_and_exports_agree():
    path = ensure_built()
    decl = declared()
    assert decl == set(cm._lib.SIGNATURES), decl ^ set(cm._lib.SIGNATURES)
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (cdcmdr_[a-z0-9_]+)\b", out))
    assert decl <= exported, decl - exported
    assert exported <= decl, f"exported but not declared in the header: {exported - decl}"


def test_library_loads_and_reports_version():
    lib = cm._lib.Lib(ensure_built())
    assert lib.version() >= 100
    assert lib.launch_count() >= 0
    assert lib.reduce_scratch_bytes() > 0 and lib.bn_scratch_bytes(64) > 0 and lib.colsum_scratch_bytes(64) > 0
    assert lib.embed_plan_bytes(1024, 1000, 16) > 0 and lib.route_scratch_bytes(1000, 4) > 0


def test_struct_sizes_match_header():
    import ctypes as C
    assert cm._lib.STEP_STATE_BYTES == 48
    assert C.sizeof(cm._lib.MixDesc) == 16 + 3 * 8 + 8          # + n_pairs (int32) padded to the pointer alignment
    assert C.sizeof(cm._lib.BnDesc) == 8 * 8 + 8 + 4 + 4 + 8 + 8
