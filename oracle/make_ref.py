#!/usr/bin/env python
"""Recipe for oracle/_ref/: a verbatim, UNTRACKED install of the reference's model package (TEST INFRASTRUCTURE, not product).

    python oracle/make_ref.py            # run HERE (the dev container), where /root/reference exists

The reference (/root/reference) is 20 loose Python files without setup.py / pyproject.toml, so `pip install --target` has
nothing to install; this script is the equivalent: it copies the files the hot path imports (model/*.py) byte for byte into
oracle/_ref/model/ and records their SHA-256 in oracle/_ref/MANIFEST.json.  oracle/_ref/ is git-ignored (the sources never
enter this repository's history) but NOT gpurun-ignored, so it travels to the GPU box with the snapshot, where
`bench.py --impl reference` and the `cpu_baseline` leg run the reference's own torch modules on the host cores
(`load_reference()` below).  Nothing in the product package imports this directory.

`__graft_entry__.build()` calls `make()` when /root/reference is present; on the GPU box the prebuilt copy is used as is.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("CDCMDR_REFERENCE", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")
FILES = ["model/layer.py", "model/ple.py", "model/mmoe.py", "model/dcn.py", "model/dcnv2.py", "model/star.py", "model/cdc.py",
         "model/pepnet.py", "model/autoint.py"]          # cdc.py imports pepnet at module scope


def make(verbose=True) -> bool:
    """Copy the reference's model files into oracle/_ref/ (outputs only there).  Returns False when /root/reference is absent."""
    if not os.path.isdir(os.path.join(REF_SRC, "model")):
        return False
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF_SRC, rel), os.path.join(REF_DST, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(REF_DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF_SRC, "files": manifest}, f, indent=1)
    if verbose:
        print(f"oracle/_ref: {len(manifest)} reference files installed from {REF_SRC}")
    return True


def available() -> bool:
    return os.path.exists(os.path.join(REF_DST, "model", "cdc.py"))


def load_reference():
    """Import the reference's model package from oracle/_ref/ (SURVEY §8c shim: matplotlib and the unshipped
    dataset.aliccp.preprocess_ali_ccp are stubbed).  Returns the `model` package namespace as a dict of classes."""
    if not available():
        raise FileNotFoundError("oracle/_ref is missing: run `python oracle/make_ref.py` where /root/reference exists")
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    ds, al = types.ModuleType("dataset"), types.ModuleType("dataset.aliccp")
    pp = types.ModuleType("dataset.aliccp.preprocess_ali_ccp")
    pp.reduce_mem = lambda df: df
    sys.modules.setdefault("dataset", ds)
    sys.modules.setdefault("dataset.aliccp", al)
    sys.modules.setdefault("dataset.aliccp.preprocess_ali_ccp", pp)
    if REF_DST not in sys.path:
        sys.path.insert(0, REF_DST)
    from model.cdc import CDC          # noqa: E402  (the reference's own modules)
    from model.ple import PLE          # noqa: E402
    from model.mmoe import MMoE        # noqa: E402
    from model.star import STAR        # noqa: E402
    return dict(CDC=CDC, PLE=PLE, MMoE=MMoE, STAR=STAR)


if __name__ == "__main__":
    ok = make()
    if not ok:
        print(f"{REF_SRC} not found: nothing installed", file=sys.stderr)
        sys.exit(1)
