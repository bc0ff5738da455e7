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


def slab_gaps(lowest, highest, slab_nz, rank):
    """z-slab EDT halo from the all-gathered column extents (include/b200_planner.h b200_map_build_edt_slab).

    lowest[r], highest[r]: int32 [Nx,Ny] lowest / highest occupied z of slab r in slab-local voxels (-1 = empty column);
    slab_nz[r]: z extent of slab r.  Returns (below_gap, above_gap) for `rank`: voxels from its first / last z-plane to
    the nearest occupied voxel in the slabs below / above (>= 1), -1 where there is none."""
    world = len(slab_nz)
    shape = np.asarray(lowest[0]).shape
    below = np.full(shape, -1, dtype=np.int32)
    gap = 0  # voxels between this slab's first plane and the top plane of slab r, minus one
    for r in range(rank - 1, -1, -1):
        hi = np.asarray(highest[r])
        cand = gap + (slab_nz[r] - hi)          # local z = hi sits (slab_nz[r] - hi) voxels below our z = 0, plus the slabs in between
        take = (hi >= 0) & (below < 0)
        below[take] = cand[take].astype(np.int32)
        gap += slab_nz[r]
    above = np.full(shape, -1, dtype=np.int32)
    gap = 0
    for r in range(rank + 1, world):
        lo = np.asarray(lowest[r])
        cand = gap + lo + 1                     # local z = lo of slab r is lo + 1 voxels above our last plane, plus the slabs in between
        take = (lo >= 0) & (above < 0)
        above[take] = cand[take].astype(np.int32)
        gap += slab_nz[r]
    return below, above


def slab_gaps_torch(lowest, highest, slab_nz, rank):
    """Same as slab_gaps on torch tensors of shape [world, Nx, Ny] (device-resident, after an NCCL all_gather)."""
    import torch
    world = lowest.shape[0]
    below = torch.full_like(lowest[0], -1)
    gap = 0
    for r in range(rank - 1, -1, -1):
        cand = gap + (slab_nz[r] - highest[r])
        take = (highest[r] >= 0) & (below < 0)
        below = torch.where(take, cand.to(below.dtype), below)
        gap += slab_nz[r]
    above = torch.full_like(lowest[0], -1)
    gap = 0
    for r in range(rank + 1, world):
        cand = gap + lowest[r] + 1
        take = (lowest[r] >= 0) & (above < 0)
        above = torch.where(take, cand.to(above.dtype), above)
        gap += slab_nz[r]
    return below.contiguous(), above.contiguous()
