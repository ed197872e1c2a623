"""Data-parallel plumbing: one process per GPU, frames sharded, one counter all-reduce.

Shard rule = the reference's eval sampler (datasets/samplers/distributed_sampler.py:29-38):
pad the index list to a multiple of the world size by wrapping around, then hand each rank ONE
CONTIGUOUS block (so that a scene's frames stay on one rank when temporal mode is on).
"""
import math
import os

import torch
import torch.distributed as dist


def contiguous_shard(num_samples_total, rank, world_size):
    """-> list of frame indices for `rank` (length ceil(n / world_size), wrap-around padded)."""
    per = int(math.ceil(num_samples_total * 1.0 / world_size))
    total = per * world_size
    indices = list(range(num_samples_total))
    indices = (indices * math.ceil(total / max(len(indices), 1)))[:total] if indices else []
    return indices[rank * per:(rank + 1) * per]


def owned_unique(num_samples_total, rank, world_size):
    """Positions (within `contiguous_shard(...)`) of the frames this rank must COUNT: the wrap-around padding
    duplicates frames 0.. on the last rank(s); the reference drops them again after collection
    (`apis/test.py:130`: results[:len(dataset)]), so a SUM all-reduce of per-rank counters must skip them."""
    per = int(math.ceil(num_samples_total * 1.0 / world_size))
    first = rank * per
    return [i for i in range(per) if first + i < num_samples_total]


def init_from_env(backend=None):
    """torchrun-style rendezvous (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).  Returns (rank, local_rank, world)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29511')
        if backend == 'nccl':
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def all_reduce_counters(vec):
    """SUM-reduce the 187 metric counters (fp64 tensor on the rank's device)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
    return vec


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
