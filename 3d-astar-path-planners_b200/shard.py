"""Multi-GPU partitioning of the hot path: the map is REPLICATED (every rank rebuilds it from the
same cloud) and a query batch is split into contiguous shards, one per rank, with NO data-path collective; results
come back with one all_gather on the control path.  Works with any torch.distributed backend (nccl on GPUs, gloo in
the CPU tests)."""
import numpy as np


def query_shard(nq, rank, world):
    """Contiguous shard [lo, hi) of nq queries for `rank` (shard i = [i*nq/P, (i+1)*nq/P))."""
    return (rank * nq) // world, ((rank + 1) * nq) // world


def find_paths_sharded(find_paths, starts, goals, rank, world, dist=None):
    """Runs `find_paths(starts[lo:hi], goals[lo:hi]) -> structured results` on this rank's shard and returns the
    full result array on every rank.  `find_paths` is Planner.findPaths-like (host arrays in, numpy structured
    array out); `dist` is torch.distributed (or None when world == 1)."""
    nq = len(starts)
    lo, hi = query_shard(nq, rank, world)
    local = find_paths(starts[lo:hi], goals[lo:hi])
    if world == 1 or dist is None:
        return local
    parts = [None] * world
    dist.all_gather_object(parts, (lo, hi, local.tobytes(), str(local.dtype.descr)))
    out = np.zeros(nq, dtype=local.dtype)
    for plo, phi, raw, _ in parts:
        out[plo:phi] = np.frombuffer(raw, dtype=local.dtype)
    return out
