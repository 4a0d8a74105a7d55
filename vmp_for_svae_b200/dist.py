"""Multi-GPU plumbing of the hot path: points shard contiguously across ranks (the reference's `tf.split` of the
minibatch across towers, data.py:174-175), global parameters are replicated, and the ONLY exchange per step is one
all-reduce(sum) of the packed [K*(D^2+D+2) + 4] double buffer (sufficient statistics + ELBO partials).
The reference instead gathers log r and x to one device (experiments.py:247-260); that O(N*K) traffic is not copied."""
import os

import torch
import torch.distributed as dist


def shard_range(n_total, rank, world_size):
    """Contiguous, balanced partition of range(n_total): the first (n_total % world) ranks get one extra point."""
    base, rem = divmod(int(n_total), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*).
    Returns (rank, world_size, local_rank); a single process without the env runs un-distributed."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29512')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def allreduce_packed(buf, group=None):
    """Sum the packed statistics buffer over ranks, in place (no-op for a single process)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


def max_over_ranks(value, device):
    """max of a python float over ranks (timing rule: device time = max over ranks)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
