"""Host-side logic of the package (arena layout, launch sequencing, backward wiring, optimizer plumbing, state_dict
compatibility) on CPU tensors: the CUDA library is replaced by the host-memory emulator of its C-ABI
(oracle/host_abi.py) and every model is driven through the SAME package code the GPU runs, against the golden fixtures
produced by the unmodified reference (tests/golden/make_golden.py).  No GPU needed."""
import numpy as np
import pytest
import torch

import cdcmdr_b200 as cm
from oracle.host_abi import HostABI
from tests.golden_cases import CASES, FIELD_DIMS, E, L2, load, state
from tests.test_oracle_golden import bias_before_bn, close
from tests.util import build_model, run_golden_case


@pytest.fixture(autouse=True)
def emulator():
    old = cm._lib._LIB
    cm._lib.install(HostABI())
    yield
    cm._lib.install(old)


SUPPORTED = sorted(n for n in CASES if build_model(n, probe=True))


@pytest.mark.parametrize("name", SUPPORTED)
def test_fused_train_step_matches_reference(name):
    run_golden_case(name, device="cpu", path="fused")


@pytest.mark.parametrize("name", SUPPORTED)
def test_autograd_path_matches_reference(name):
    run_golden_case(name, device="cpu", path="autograd")


def test_state_dict_keys_match_reference():
    for name in SUPPORTED:
        gold = load(name)
        model = build_model(name)
        ref = {k[len("sd0."):]: v.shape for k, v in gold.items() if k.startswith("sd0.")}
        ours = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        assert set(ours) == set(ref), (name, set(ours) ^ set(ref))
        for k in ref:
            assert ours[k] == tuple(ref[k]), (name, k)


def test_real_library_refuses_cpu_tensors():
    cm._lib.install(None)
    cm._lib._LIB = None
    model = build_model("ple")
    with pytest.raises(RuntimeError, match="CUDA device"):
        model(torch.zeros(4, len(FIELD_DIMS), dtype=torch.int32))
