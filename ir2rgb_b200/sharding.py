"""Batch sharding of the flow hot path across GPUs.

Frame pairs are independent for every operator on the path and for FlowNet2 inference, so the only
multi-GPU structure is a partition of the batch: rank r of N owns one contiguous chunk, weights are
replicated, and there is NO collective or exchange step on the data path (NCCL is used by bench.py only
for the barrier and the max-over-ranks of the timing).
"""


def shard_bounds(n_items, rank, world):
    """[start, stop) of rank's contiguous chunk; chunks differ by at most one item and tile [0, n_items)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard(tensor, rank, world):
    """This rank's contiguous batch chunk of `tensor` (a view, no copy)."""
    s, e = shard_bounds(tensor.shape[0], rank, world)
    return tensor[s:e]
