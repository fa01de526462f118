"""POI sharding across GPUs (one process per GPU, torch.distributed for the plumbing).

The path shards trivially: POIs are independent in both stages
(reference src/oc_fftcc.cpp:280-284, src/oc_icgn.cpp:346-350), the images are read-only.  The only
exchanges are a broadcast of the image/volume pair from rank 0 before the compute and a gather of
the POI records to rank 0 after it; there is no collective inside the hot path.

Works with backend "nccl" (device tensors, NVLink) and "gloo" (CPU tensors; used by the tests).
"""
import torch
import torch.distributed as dist


def shard_bounds(n, world_size, rank):
    """Contiguous block split [lo, hi) of n POIs; earlier ranks take the remainder.  Contiguous
    blocks keep spatial locality (neighbouring POIs share image tiles in L2)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad world_size/rank")
    base, rem = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def all_shard_sizes(n, world_size):
    return [shard_bounds(n, world_size, r)[1] - shard_bounds(n, world_size, r)[0] for r in range(world_size)]


def broadcast_images(ref, tar, src=0):
    """In-place broadcast of the image (or volume) pair from `src` to every rank."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(ref, src=src)
        dist.broadcast(tar, src=src)
    return ref, tar


def scatter_pois(all_pois, n_total, floats, device, src=0):
    """Rank `src` holds the full [n_total, floats] queue; every rank returns its own shard.
    (Ranks that already hold the queue can simply slice with shard_bounds.)"""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_bounds(n_total, world, rank)
    mine = torch.empty((hi - lo, floats), dtype=torch.float32, device=device)
    if world == 1:
        mine.copy_(all_pois[lo:hi])
        return mine
    if rank == src:
        chunks = [all_pois[slice(*shard_bounds(n_total, world, r))].contiguous() for r in range(world)]
        mine.copy_(chunks[src])
        reqs = [dist.isend(chunks[r], dst=r) for r in range(world) if r != src]
        for q in reqs:
            q.wait()
    else:
        dist.recv(mine, src=src)
    return mine


def gather_pois(shard, n_total, dst=0):
    """Gather the per-rank shards back into one [n_total, floats] tensor on rank `dst`
    (returns None elsewhere).  Shards may differ in length by one record."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if world == 1:
        return shard
    floats = shard.shape[1]
    sizes = all_shard_sizes(n_total, world)
    if rank == dst:
        out = torch.empty((n_total, floats), dtype=shard.dtype, device=shard.device)
        lo, hi = shard_bounds(n_total, world, dst)
        out[lo:hi].copy_(shard)
        reqs = []
        for r in range(world):
            if r == dst or sizes[r] == 0:
                continue
            lo, hi = shard_bounds(n_total, world, r)
            reqs.append(dist.irecv(out[lo:hi], src=r))
        for q in reqs:
            q.wait()
        return out
    if shard.shape[0]:
        dist.send(shard.contiguous(), dst=dst)
    return None
