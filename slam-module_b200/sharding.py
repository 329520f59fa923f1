"""Multi-GPU partitioning of the hot path: frames and keyframe pairs are independent, so a
job is split into contiguous blocks by rank, every rank runs the same batched kernels on its block with
its own context, and the only communication is the final host gather.  No collective touches the data
path; torch.distributed (NCCL on GPUs, gloo in the CPU tests) carries the barrier, the max-over-ranks of
the device time and the gather of the small result arrays.
"""
import numpy as np


def block_range(n, rank, world):
    """Balanced contiguous block of range(n) owned by `rank`: the first n % world ranks get one extra."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("rank %d outside world %d" % (rank, world))
    base, extra = divmod(int(n), world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def round_robin(n, rank, world):
    """Frame i -> rank i % world (streams: keeps every GPU busy from the first frame on)."""
    return np.arange(rank, n, world)


def n_unordered_pairs(n_sets):
    return n_sets * (n_sets - 1) // 2


def unordered_pair(k, n_sets):
    """k-th pair (i < j) of the row-major enumeration (0,1), (0,2), ..., (n-2,n-1); vectorised, exact in int64."""
    k = np.asarray(k, np.int64)
    n = np.int64(n_sets)
    # row i starts at offset i*n - i*(i+1)/2 ; invert with a float guess and fix up exactly
    i = np.floor(((2 * n - 1) - np.sqrt(np.maximum((2 * n - 1) ** 2 - 8.0 * k, 0.0))) / 2).astype(np.int64)
    start = lambda r: r * n - r * (r + 1) // 2
    i = np.where(start(i) > k, i - 1, i)
    i = np.where(start(i + 1) <= k, i + 1, i)
    j = k - start(i) + i + 1
    return i, j


def pair_block(n_sets, rank, world, limit=None):
    """The block of the all-pairs candidate list (config 5: every unordered keyframe pair) owned by `rank`,
    as an int32 [m, 2] array, generated from the linear pair index so no rank materialises the whole list.
    `limit` truncates the block (bounded bench samples)."""
    lo, hi = block_range(n_unordered_pairs(n_sets), rank, world)
    if limit is not None:
        hi = min(hi, lo + int(limit))
    i, j = unordered_pair(np.arange(lo, hi, dtype=np.int64), n_sets)
    return np.stack([i, j], axis=1).astype(np.int32)


def host_gather(local, td=None, dst=0):
    """Final host gather: every rank contributes a numpy array (first dimension may differ), rank `dst` gets
    the concatenation in rank order (others get None).  Without a process group: identity."""
    local = np.ascontiguousarray(local)
    if td is None or not td.is_initialized() or td.get_world_size() == 1:
        return local
    world, rank = td.get_world_size(), td.get_rank()
    parts = [None] * world if rank == dst else None
    td.gather_object(local, parts, dst=dst)
    return np.concatenate(parts, axis=0) if rank == dst else None


def max_over_ranks(value, td=None, device=None):
    """max over ranks of a python float (the bench's device time)."""
    if td is None or not td.is_initialized() or td.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t.item())
