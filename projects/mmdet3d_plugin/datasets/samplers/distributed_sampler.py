"""Evaluation sampler with the reference's contiguous-block shard rule (one block per rank, wrap-around padding)."""
import math

import torch
from torch.utils.data import DistributedSampler as _DistributedSampler

from occnet_b200.dist import contiguous_shard


class DistributedSampler(_DistributedSampler):
    def __init__(self, dataset=None, num_replicas=None, rank=None, shuffle=True, seed=0):
        super().__init__(dataset, num_replicas=num_replicas, rank=rank, shuffle=shuffle)
        self.seed = seed if seed is not None else 0

    def __iter__(self):
        if self.shuffle:
            raise AssertionError('shuffle is not supported by the evaluation sampler')
        idx = contiguous_shard(len(self.dataset), self.rank, self.num_replicas)
        assert len(idx) == int(math.ceil(len(self.dataset) / self.num_replicas))
        return iter(idx)
