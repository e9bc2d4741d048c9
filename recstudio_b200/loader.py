"""Device-resident batch producer.

The reference builds every training batch on the host: ``DataSampler`` draws a CPU ``randperm`` and
yields index chunks, ``DataLoader(batch_size=None)`` indexes the ``TensorFrame`` per chunk, and the
trainer copies the dict to the device (recstudio/data/dataset.py:1083-1123,1687-1734,
recommender.py:596,699-714).  At > 4 M interactions/s (hundreds of steps per second) that Python/H2D
loop is the limiter.  ``DeviceBatchLoader`` keeps the interaction columns on the GPU and slices
batches there; it goes behind the unchanged ``train_loader()`` / ``_get_train_loaders()`` surface.

The epoch permutation is drawn EXACTLY like ``DataSampler.__iter__`` (a fresh CPU generator seeded from
the global CPU RNG, ``torch.randperm(n, generator=...)``), so a run sees the same batches as the
reference trainer with the same seed; only the permutation (8 B per interaction per epoch) crosses
PCIe.  Plumbing only: tensors are indexed with torch on the device, no arithmetic of the path lives here.
"""
from __future__ import annotations

import torch


class DeviceBatchLoader:
    def __init__(self, columns: dict, batch_size: int, device, shuffle: bool = True, drop_last: bool = False, dataset=None):
        self.device = torch.device(device)
        self.columns = {k: v.to(self.device) for k, v in columns.items() if isinstance(v, torch.Tensor)}
        n = {v.shape[0] for v in self.columns.values()}
        if len(n) != 1:
            raise ValueError("all columns must have the same number of rows")
        self.n = n.pop()
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), shuffle, drop_last
        self.dataset = dataset

    @classmethod
    def from_dataset(cls, train_data, batch_size: int, device, shuffle: bool = True, drop_last: bool = False):
        """Materialise the training split once: ``train_data[arange(len)]`` is the dict the reference's
        loader would deliver row by row (dataset.py:893-913)."""
        if getattr(train_data, "data_index", torch.zeros(1)).dim() > 1:
            raise ValueError("DeviceBatchLoader covers TripletDataset-style splits (1-D data_index); sequence datasets "
                             "use the reference's SortedDataSampler")
        train_data.eval_mode = False                              # what train_loader() does first (dataset.py:1103)
        full = train_data[torch.arange(len(train_data))]
        return cls({k: v for k, v in full.items() if isinstance(v, torch.Tensor) and v.dim() >= 1 and v.shape[0] == len(train_data)},
                   batch_size, device, shuffle, drop_last, dataset=train_data)

    def __len__(self):
        return self.n // self.batch_size if self.drop_last else (self.n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        # torch's DataLoader iterator draws its `_base_seed` from the global CPU generator BEFORE the (lazy)
        # DataSampler body runs (torch/utils/data/dataloader.py:705-710): consume the same number first
        torch.empty((), dtype=torch.int64).random_()
        if self.shuffle:                                         # dataset.py:1710-1722, same RNG consumption
            generator = torch.Generator()
            generator.manual_seed(int(torch.empty((), dtype=torch.int64).random_().item()))
            order = torch.randperm(self.n, generator=generator).to(self.device, non_blocking=True)
        else:
            order = torch.arange(self.n, device=self.device)
        chunks = order.split(self.batch_size)
        if self.drop_last and len(chunks[-1]) < self.batch_size:
            chunks = chunks[:-1]
        for idx in chunks:
            yield {k: v[idx] for k, v in self.columns.items()}
