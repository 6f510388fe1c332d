"""Multi-GPU sharding of a batch of independent streams: contiguous ranges of the stream index
balanced by compressed bytes; no data-path collective -- ranks only agree on the job totals."""
import numpy as np


def partition_by_bytes(lengths, n_parts):
    """[(begin, end)] * n_parts: contiguous index ranges whose byte sums are as even as a prefix-sum split allows."""
    lengths = np.asarray(lengths, dtype=np.uint64)
    n = lengths.size
    if n == 0:
        return [(0, 0)] * n_parts
    csum = np.concatenate([[0], np.cumsum(lengths.astype(np.float64))])
    total = csum[-1]
    cuts = [0]
    for k in range(1, n_parts):
        target = total * k / n_parts
        i = int(np.searchsorted(csum, target, side="left"))
        # pick the closer of the two neighbouring cut points
        if i > 0 and abs(csum[i - 1] - target) <= abs(csum[min(i, n)] - target):
            i -= 1
        cuts.append(max(cuts[-1], min(i, n)))
    cuts.append(n)
    return [(cuts[k], cuts[k + 1]) for k in range(n_parts)]


def reduce_job(units_done, seconds):
    """All-reduce the per-rank (units, time) into (sum of units, max time) -- the only collective of a sharded job."""
    import torch.distributed as dist
    u = units_done.clone()
    t = seconds.clone()
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(u.item()), float(t.item())
