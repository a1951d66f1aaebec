"""GPU-resident ray table and on-device batch sampler.

Replaces the step before the path: the reference keeps every training ray in one host tensor per key
(`BaseRaysDataset._combine`, framework/datasets.py:238-266), serves single rays through
`Dataset.__getitem__` (baseline/dataset/satnerf_dataset.py:122-133), collates `batch_size` of them per
step in DataLoader workers and copies the batch to the GPU (framework/pipelines.py:107-118:
`DataLoader(shuffle=cfgs.run.shuffle_dataset, batch_size=cfgs.pipeline.batch_size, pin_memory=True)`, one
loader per dataset key, combined by Lightning).  At > 10^5 rays/s the Python collate is the bottleneck, so
here the combined table lives on the device once and a batch is one `index_select` per key - no host work,
no H2D copy, no sync.

Semantics kept from the reference: one pass over every ray per epoch in a fresh random order
(`shuffle=True`; a seeded generator makes the order reproducible), the last batch of an epoch may be short
(`drop_last=False`), several tables (colour rays / depth-supervision rays) are stepped together and the
shorter ones restart (Lightning's `max_size_cycle`).  Data-parallel ranks take disjoint contiguous shards of
every global batch (`dist.shard_range`), so the union over ranks is exactly the single-process batch.
"""
from __future__ import annotations

from typing import Dict, Iterator, Optional

import torch

from .dist import shard_range


class DeviceRayTable:
    """All rays of one dataset key (`combined_data`), resident on `device`."""

    def __init__(self, tensors: Dict[str, torch.Tensor], device="cuda"):
        n = {int(v.shape[0]) for v in tensors.values()}
        if len(n) != 1:
            raise ValueError(f"every tensor of a ray table needs the same first dimension, got {sorted(n)}")
        self.n_rays = n.pop()
        self.device = torch.device(device)
        self.tensors = {k: v.to(self.device).contiguous() for k, v in tensors.items()}

    @classmethod
    def from_items(cls, items, device="cuda"):
        """`items`: the reference's list of per-image dicts; tensors are concatenated along dim 0 as
        `_combine` does (framework/datasets.py:238-266), non-tensor entries are skipped."""
        keys = [k for k, v in items[0].items() if torch.is_tensor(v)]
        return cls({k: torch.cat([it[k] for it in items], 0) for k in keys}, device)

    def __len__(self):
        return self.n_rays

    def gather(self, idx: torch.Tensor) -> Dict[str, torch.Tensor]:
        return {k: v.index_select(0, idx) for k, v in self.tensors.items()}

    def epoch(self, batch_size: int, shuffle: bool = True, seed: int = 0, epoch: int = 0, rank: int = 0,
              world: int = 1, drop_last: bool = False) -> Iterator[Dict[str, torch.Tensor]]:
        """Batches of one epoch.  `batch_size` is the GLOBAL batch; this rank receives its shard of each.
        The permutation is drawn on the device from (seed, epoch): every rank draws the same one."""
        if shuffle:
            g = torch.Generator(device=self.device)
            g.manual_seed((seed * 1_000_003 + epoch) & 0x7FFFFFFFFFFFFFFF)
            perm = torch.randperm(self.n_rays, device=self.device, generator=g)
        else:
            perm = torch.arange(self.n_rays, device=self.device)
        for lo in range(0, self.n_rays, batch_size):
            hi = min(lo + batch_size, self.n_rays)
            if drop_last and hi - lo < batch_size:
                return
            a, b = shard_range(hi - lo, rank, world)
            out = self.gather(perm[lo + a:lo + b])
            # what a data-parallel step needs besides its shard: the size of the global batch (the loss means run over
            # it) and the shard's position in it (the Philox key of a ray is its index in the global batch)
            out["_global_rays"] = hi - lo
            out["_ray_offset"] = a
            yield out

    def validate_labels(self, n_classes: int, key: str = "semantic", ignore_index: int = -100):
        """Every label must be a class index in [0, n_classes) (or ignore_index): torch's CrossEntropyLoss raises for
        anything else, the fused loss kernel would silently skip such rays.  One device read, once per table."""
        from .autograd import as_labels, label_counts
        c = label_counts(as_labels(self.tensors[key]), None, n_classes, ignore_index, -1)
        bad = int(c[2].item())
        if bad:
            raise ValueError(f"{bad} labels of '{key}' lie outside [0, {n_classes}): class-count mismatch between the "
                             f"dataset and the model's semantic_n_classes?")

    def steps_per_epoch(self, batch_size: int, drop_last: bool = False) -> int:
        return self.n_rays // batch_size if drop_last else -(-self.n_rays // batch_size)


class ZippedTables:
    """Several tables stepped together, the shorter ones restarting until the longest has finished one epoch
    (how Lightning combines the dict of train loaders: framework/pipelines.py:107-118)."""

    def __init__(self, tables: Dict[str, DeviceRayTable], batch_sizes: Dict[str, int], seed: int = 0, rank: int = 0,
                 world: int = 1):
        self.tables, self.batch_sizes, self.seed, self.rank, self.world = tables, batch_sizes, seed, rank, world

    def epoch(self, epoch: int = 0) -> Iterator[Dict[str, Optional[Dict[str, torch.Tensor]]]]:
        steps = max(t.steps_per_epoch(self.batch_sizes[k]) for k, t in self.tables.items())
        its, restarts = {}, {k: 0 for k in self.tables}

        def start(k):
            # a restarted (shorter) table gets a fresh order each time round
            return self.tables[k].epoch(self.batch_sizes[k], True, self.seed, epoch * 1009 + restarts[k], self.rank, self.world)

        for k in self.tables:
            its[k] = start(k)
        for _ in range(steps):
            out = {}
            for k in self.tables:
                try:
                    out[k] = next(its[k])
                except StopIteration:
                    restarts[k] += 1
                    its[k] = start(k)
                    out[k] = next(its[k])
            yield out
