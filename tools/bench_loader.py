"""GPU timing of batch assembly (not a test, not the bench contract): cdcmdr_b200.DeviceLoader against the loader
the reference builds, DataLoader(TensorDataset(X, y, group), bs, shuffle=True) over device-resident tensors (run.py:240-244).

    python tools/bench_loader.py [--rows 2000000]

Per batch size: ms per batch of each loader (torch's loader indexes the device tensors one sample at a time, so only a few of its
batches are timed) and the bytes DeviceLoader moves per second (read + write of X int32 [B, 23], y int16 [B, 1], group int64
[B, 1]).  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
from torch.utils.data import DataLoader, TensorDataset

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cdcmdr_b200 as cm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=2_000_000)
    a = ap.parse_args()
    rng = np.random.default_rng(0)
    X = torch.from_numpy(rng.integers(0, 45_000, size=(a.rows, 23)).astype(np.int32)).cuda()
    y = torch.from_numpy((rng.random((a.rows, 1)) < 0.05).astype(np.int16)).cuda()
    g = torch.from_numpy(rng.integers(0, 4, size=(a.rows, 1)).astype(np.int64)).cuda()
    ds = TensorDataset(X, y, g)
    out = dict(what="batch assembly over device-resident tensors", rows=a.rows, cases=[])
    for bs, n_ref in ((2048, 3), (65536, 1)):
        torch.manual_seed(1)
        it = iter(DataLoader(ds, bs, shuffle=True))
        first_ref = next(it)                                             # includes the epoch's randperm
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_ref):
            next(it)
        torch.cuda.synchronize()
        ref_ms = 1e3 * (time.perf_counter() - t0) / n_ref
        torch.manual_seed(1)
        ld = cm.DeviceLoader(ds, bs, shuffle=True)
        it = iter(ld)
        first = next(it)
        same = all(torch.equal(p, q) for p, q in zip(first_ref, first))
        n_mine = min(len(ld) - 2, 200)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_mine):
            next(it)
        torch.cuda.synchronize()
        ms = 1e3 * (time.perf_counter() - t0) / n_mine
        t0 = time.perf_counter()
        for _ in ld:                                                     # a whole epoch, permutation + upload included
            pass
        torch.cuda.synchronize()
        epoch_ms = 1e3 * (time.perf_counter() - t0)
        moved = 2 * bs * (23 * 4 + 2 + 8)
        out["cases"].append(dict(batch=bs, torch_dataloader_ms_per_batch=ref_ms, device_loader_ms_per_batch=ms,
                                 device_loader_epoch_ms=epoch_ms, batches_per_epoch=len(ld), first_batch_identical=bool(same),
                                 device_loader_gb_per_s=moved / (ms * 1e-3) / 1e9))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
