"""Batch assembly: cdcmdr_b200.DeviceLoader against the loader the reference builds,
`DataLoader(TensorDataset(X, y[, group]), bs, shuffle=True)` (run.py:240-244, 272-274) - same batches, same order, bit for bit,
same consumption of torch's global RNG (so everything seeded downstream of the loaders stays in step too)."""
import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader, TensorDataset

import cdcmdr_b200 as cm
from oracle.host_abi import HostABI


def _dataset(n, device, with_group=True, seed=0):
    rng = np.random.default_rng(seed)
    X = torch.from_numpy(rng.integers(0, 1000, size=(n, 23)).astype(np.int32)).to(device)      # run.py:198 ids int32
    y = torch.from_numpy((rng.random((n, 1)) < 0.3).astype(np.int16)).to(device)               # run.py:199 labels int16
    g = torch.from_numpy(rng.integers(0, 4, size=(n, 1)).astype(np.int64)).to(device)          # run.py:230 group int64
    return TensorDataset(X, y, g) if with_group else TensorDataset(X, y)


def _check(device):
    # 1. same seed -> same shuffles, epoch after epoch; ragged last batch; drop_last; sequential
    for n, bs, kw in ((1000, 64, dict(shuffle=True)), (257, 256, dict(shuffle=True)), (130, 32, dict(shuffle=True, drop_last=True)),
                      (77, 16, dict(shuffle=False)), (5, 8, dict(shuffle=True)), (64, 64, dict(shuffle=True))):
        ds = _dataset(n, device, with_group=(n % 2 == 0))
        torch.manual_seed(2000)
        ref = [list(DataLoader(ds, bs, **kw)) for _ in range(2)]
        after_ref = torch.rand(3)
        torch.manual_seed(2000)
        ld = cm.DeviceLoader(ds, bs, **kw)
        mine = [list(ld) for _ in range(2)]
        after_mine = torch.rand(3)
        assert len(ld) == len(DataLoader(ds, bs, **kw))
        assert torch.equal(after_ref, after_mine), "global RNG consumed differently"
        for ea, eb in zip(ref, mine):
            assert len(ea) == len(eb) == len(ld)
            for ba, bb in zip(ea, eb):
                for ta, tb in zip(ba, bb):
                    assert ta.dtype == tb.dtype and ta.shape == tb.shape and ta.device == tb.device and torch.equal(ta, tb)
    # 2. an explicit generator
    ds = _dataset(300, device)
    g1, g2 = torch.Generator().manual_seed(9), torch.Generator().manual_seed(9)
    for ba, bb in zip(DataLoader(ds, 50, shuffle=True, generator=g1), cm.DeviceLoader(ds, 50, shuffle=True, generator=g2)):
        assert all(torch.equal(ta, tb) for ta, tb in zip(ba, bb))
    # 3. the per-domain loaders of CDC (run.py:265-274) driven like get_domain_data (run.py:499-505): interleaved iterators that
    #    restart when exhausted
    X, y, _ = _dataset(600, device).tensors
    dom = (X[:, 10] % 3)

    def loaders(cls):
        return [cls(TensorDataset(X[dom == d], y[dom == d]), 64, shuffle=True) for d in range(3)]

    def drive(lds, seq):
        its = [iter(ld) for ld in lds]
        out = []
        for d in seq:
            try:
                out.append(next(its[d]))
            except StopIteration:
                its[d] = iter(lds[d])
                out.append(next(its[d]))
        return out
    seq = np.random.default_rng(1).integers(0, 3, size=40).tolist()
    torch.manual_seed(7)
    a = drive(loaders(DataLoader), seq)
    torch.manual_seed(7)
    b = drive(loaders(cm.DeviceLoader), seq)
    for ba, bb in zip(a, b):
        assert all(torch.equal(ta, tb) for ta, tb in zip(ba, bb))


@pytest.fixture
def emulator():
    old = cm._lib._LIB
    cm._lib.install(HostABI())
    yield
    cm._lib.install(old)


def test_device_loader_matches_torch_dataloader_host_logic(emulator):
    _check("cpu")


def test_device_loader_refuses_host_tensors_and_odd_dtypes(emulator):
    cm._lib.install(None)                                      # the real library: no emulator flag
    ds = _dataset(10, "cpu")
    try:
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            cm.DeviceLoader(ds, 4)
    finally:
        cm._lib.install(HostABI())
    with pytest.raises(TypeError):
        cm.DeviceLoader(TensorDataset(torch.zeros(4, 2, dtype=torch.uint8)), 2)
    with pytest.raises(ValueError):
        cm.DeviceLoader((torch.zeros(4, 2), torch.zeros(5, 2)), 2)


@pytest.mark.gpu
def test_device_loader_matches_torch_dataloader_gpu():
    _check("cuda")
