"""Data-parallel plumbing.  Sequences are independent: the batch axis shards across
ranks with all K particles of a sequence on one rank, parameters replicated, and NO collective on the
data path of the forward pass.  The only cross-rank step is combining the per-rank objective
scalars (batch means) -- one tiny all-reduce.  Noise is keyed by the global row (`row_offset`), so
a sharded run reproduces the unsharded draws."""
import torch
import torch.distributed as dist


def shard_range(n_items, world_size, rank):
    """Contiguous, balanced split of `n_items` sequences: (start, count) for `rank`."""
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def row_offset(n_items, world_size, rank, k_particles):
    """Global index of this rank's first row (row = b*K + k)."""
    return shard_range(n_items, world_size, rank)[0] * k_particles


def combine_batch_means(local_means, n_local, group=None):
    """Batch means computed per rank -> global batch means: sum_r(mean_r * n_r) / sum_r n_r (one all-reduce)."""
    buf = torch.cat([local_means.reshape(-1).double() * float(n_local),
                     torch.tensor([float(n_local)], dtype=torch.float64, device=local_means.device)])
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return (buf[:-1] / buf[-1]).to(local_means.dtype).reshape(local_means.shape)


def world_size(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def allreduce_flat_gradient(flat_grad, n_local, n_global, group=None):
    """The one collective of a training step: every rank holds the gradient of ITS shard's batch mean in
    the flat parameter layout (`sqair_param_layout` order, from `sqair_backward`); the global-batch gradient is
    sum_r grad_r * n_r / n_global.  In place, one all-reduce (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if n_local != n_global:
            flat_grad.mul_(float(n_local) / float(n_global))
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    elif n_local != n_global:
        flat_grad.mul_(float(n_local) / float(n_global))
    return flat_grad


def broadcast_parameters(store, src=0, group=None):
    """Replicas start from rank `src`'s variables (data-parallel training keeps them identical afterwards because every
    rank applies the same all-reduced gradient)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(store.flat, src=src, group=group)
        store.mark_dirty()
    return store
