"""Batch assembly for device-resident datasets: the drop-in for

    DataLoader(TensorDataset(X, y[, group]), batch_size, shuffle=True)          run.py:240-244, 272-274

when the tensors already live on the GPU, which is how the reference builds every loader (the whole dataset is uploaded once,
run.py:239, 271).  torch's DataLoader then indexes the device tensors ONE SAMPLE AT A TIME from Python and stacks the pieces
(TensorDataset has no batched __getitems__): 3 x batch_size tiny kernels per batch, 36 ms for a 2 048-row batch on the CPU path and
far longer than the training step at 65 536 rows.  DeviceLoader yields THE SAME batches in THE SAME order:

  * the epoch's sample order is drawn exactly the way DataLoader + RandomSampler draw it (torch/utils/data/dataloader.py,
    sampler.py): creating the iterator takes one int64 from the loader's generator (the global CPU generator if none) for the
    worker base seed; the first batch then makes the permutation - `torch.randperm(n, generator=g)` where g is the loader's
    generator or, without one, a fresh Generator seeded with another int64 from the global generator.  Under the same
    `torch.manual_seed` both loaders therefore consume the global RNG identically and shuffle identically
    (tests/test_device_loader.py compares them batch for batch, bit for bit);
  * the permutation goes to the device once per epoch (int32), and each batch is one row gather per tensor
    (`cdcmdr_permute_rows`, dst[i, :] = src[perm[i], :]) into fresh tensors - byte moves, bit-exact.

torch supplies the generator and device memory; no torch operator touches the data."""
from __future__ import annotations

import torch

from . import _lib

_ELT = {1: None, 2: 2, 4: 4, 8: 8}


class DeviceLoader:
    def __init__(self, dataset, batch_size=1, shuffle=False, drop_last=False, generator=None):
        tensors = tuple(dataset.tensors) if hasattr(dataset, "tensors") else tuple(dataset)
        if not tensors:
            raise ValueError("DeviceLoader needs at least one tensor")
        n = int(tensors[0].shape[0])
        dev = tensors[0].device
        for t in tensors:
            if t.dim() < 1 or int(t.shape[0]) != n:
                raise ValueError("Size mismatch between tensors")                   # TensorDataset's own check
            if t.device != dev:
                raise ValueError("DeviceLoader: all tensors must live on one device")
            if _ELT.get(t.element_size()) is None:
                raise TypeError(f"DeviceLoader: {t.dtype} rows are not supported (2-, 4- and 8-byte elements; the reference "
                                "holds int32 ids, int16 labels and int64 groups, run.py:198-199, 229-230)")
        if n >= 2 ** 31:
            raise ValueError("DeviceLoader: more than 2^31-1 samples")
        if batch_size is None or int(batch_size) <= 0:
            raise ValueError("batch_size should be a positive integer value")
        self.lib = _lib.load()
        emu = getattr(self.lib, "is_host_emulator", False)
        if dev.type != "cuda" and not emu:
            raise RuntimeError("cdcmdr: DeviceLoader gathers batches on a CUDA device (sm_100a); there is no CPU fallback - keep "
                               "torch.utils.data.DataLoader for host tensors")
        self.tensors = tuple(t.contiguous() for t in tensors)
        self.dataset = dataset
        self.n, self.device = n, dev
        self.batch_size, self.shuffle, self.drop_last, self.generator = int(batch_size), bool(shuffle), bool(drop_last), generator

    def __len__(self):
        if self.drop_last:
            return self.n // self.batch_size
        return (self.n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        return _Epoch(self)


class _Epoch:
    def __init__(self, loader: DeviceLoader):
        self.ld = loader
        # _BaseDataLoaderIter.__init__: the worker base seed is drawn even with num_workers=0
        torch.empty((), dtype=torch.int64).random_(generator=loader.generator)
        self.perm = None
        self.pos = 0
        self.left = len(loader)

    def __iter__(self):
        return self

    def __len__(self):
        return len(self.ld)

    def _order(self):
        ld = self.ld
        if not ld.shuffle:
            return None                                                              # SequentialSampler: identity
        g = ld.generator
        if g is None:                                                                # RandomSampler.__iter__
            seed = int(torch.empty((), dtype=torch.int64).random_().item())
            g = torch.Generator()
            g.manual_seed(seed)
        return torch.randperm(ld.n, generator=g).to(torch.int32).to(ld.device)

    def __next__(self):
        ld = self.ld
        if self.left <= 0:
            raise StopIteration
        if self.pos == 0:
            self.perm = self._order()
        b = min(ld.batch_size, ld.n - self.pos)
        stream = torch.cuda.current_stream(ld.device).cuda_stream if ld.device.type == "cuda" else 0
        out = []
        for t in ld.tensors:
            cols = 1
            for s in t.shape[1:]:
                cols *= int(s)
            dst = torch.empty((b,) + tuple(t.shape[1:]), dtype=t.dtype, device=ld.device)
            if cols:
                es = t.element_size()
                if self.perm is None:
                    ld.lib.permute_rows(t.data_ptr() + self.pos * cols * es, cols, None, b, cols, es, dst.data_ptr(), cols, 0, stream)
                else:
                    ld.lib.permute_rows(t.data_ptr(), cols, self.perm.data_ptr() + 4 * self.pos, b, cols, es, dst.data_ptr(), cols, 0,
                                        stream)
            out.append(dst)
        self.pos += b
        self.left -= 1
        return out
