"""Multi-GPU plumbing: streams are independent (reference: no globals, read-only weights — SURVEY 8e), so a batch is
sharded by contiguous stream ranges, one process per GPU, with NO collective on the data path.  The only exchange is
the final PCM gather to rank 0 (NCCL over NVLink on GPUs; the same code runs over gloo on CPU tensors in the tests)."""


def shard_range(n_total, rank, world):
    """Contiguous stream range [lo, hi) of `rank`; sizes differ by at most one, earlier ranks take the remainder."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_pcm(local_pcm, n_total, dist, dst=0):
    """Gather per-rank PCM shards [n_local][T] (torch int16 tensors, CPU or CUDA) to `dst` as [n_total][T].
    Shards may be ragged (see shard_range): they are padded to the largest shard for the collective."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    T = local_pcm.shape[1]
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx, T), dtype=local_pcm.dtype, device=local_pcm.device)
    pad[: local_pcm.shape[0]] = local_pcm
    raw = pad.view(torch.uint8)                       # int16 is not a collective dtype on every backend (gloo): move bytes
    outs = [torch.empty_like(raw) for _ in range(world)] if rank == dst else None
    dist.gather(raw, outs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([outs[r].view(local_pcm.dtype)[: hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)
