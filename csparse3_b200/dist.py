"""Batch sharding across the GPUs of one box.

Systems of a same-pattern batch are independent: rank g of G owns the contiguous slice
[g*ceil(B/G), (g+1)*ceil(B/G)) and refactors + solves it with NO inter-GPU traffic; the symbolic object is
replicated.  The only collective is the final gather of the solution shards (NCCL all-gather over
NVLink/NVSwitch on GPU tensors; gloo on CPU tensors in the host-logic tests)."""
import torch
import torch.distributed as dist


def shard_range(batch, rank, world):
    per = -(-batch // world)
    start = min(rank * per, batch)
    return start, min(start + per, batch)


def gather_solutions(x_local, batch, group=None):
    """All-gather the per-rank solution shards [count_r, n] into the full [batch, n] tensor (every rank gets
    it).  Shards are padded to ceil(batch/world) rows so a single all_gather_into_tensor suffices."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = -(-batch // world)
    n = x_local.shape[-1]
    start, stop = shard_range(batch, rank, world)
    assert x_local.shape[0] == stop - start
    send = x_local
    if stop - start != per:
        send = torch.zeros((per, n), dtype=x_local.dtype, device=x_local.device)
        send[:stop - start] = x_local
    out = torch.empty((world * per, n), dtype=x_local.dtype, device=x_local.device)
    dist.all_gather_into_tensor(out, send.contiguous(), group=group)
    return out[:batch]


def gather_to_root(x_local, batch, root=0, group=None):
    """Gather the solution shards on `root` only (what a consumer that post-processes on one rank needs: 1/world of
    the all-gather's receive traffic on the other ranks).  Returns the full [batch, n] tensor on root, None elsewhere.
    Shards are padded to ceil(batch/world) rows."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = -(-batch // world)
    n = x_local.shape[-1]
    start, stop = shard_range(batch, rank, world)
    assert x_local.shape[0] == stop - start
    send = x_local
    if stop - start != per:
        send = torch.zeros((per, n), dtype=x_local.dtype, device=x_local.device)
        send[:stop - start] = x_local
    if rank == root:
        out = torch.empty((world * per, n), dtype=x_local.dtype, device=x_local.device)
        dist.gather(send.contiguous(), list(out.view(world, per, n).unbind(0)), dst=root, group=group)
        return out[:batch]
    dist.gather(send.contiguous(), None, dst=root, group=group)
    return None
