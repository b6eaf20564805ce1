"""Frame sharding across the GPUs of one box (one process per GPU, torch.distributed).

The hot path is embarrassingly parallel -- every output frame depends on one source frame and one
3x3 -- so ranks share nothing and the only collective is the gather of BEV outputs / boxes to
rank 0 (NCCL over NVLink; gloo in the CPU tests).  SURVEY.md 8e.
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous [begin, end) slice of n_items owned by `rank`; sizes differ by at most one."""
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_cameras(n_cameras, rank, world):
    """Camera streams owned by `rank` (BASELINE configs[3]: 8 streams on 2/4/8 GPUs)."""
    b, e = shard_range(n_cameras, rank, world)
    return list(range(b, e))


def gather_to_rank0(t, group=None, chunks=4):
    """Gather equally-shaped per-rank tensors on rank 0 (others get None), first dim concatenated
    in rank order.  The transfer is issued in `chunks` pieces so a producer can overlap it with the
    next warp chunk; rank 0's NVLink ingress (~770 GB/s measured) is the bound, not the warp."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return t
    rank = dist.get_rank(group)
    t = t.contiguous()
    out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device) if rank == 0 else None
    n = t.shape[0]
    chunks = max(1, min(chunks, n))
    for c in range(chunks):
        b, e = shard_range(n, c, chunks)
        if e <= b:
            continue
        piece = t[b:e]
        dst_list = [out[r, b:e] for r in range(world)] if rank == 0 else None
        if rank == 0 and not all(d.is_contiguous() for d in dst_list):
            tmp = [torch.empty_like(piece) for _ in range(world)]
            dist.gather(piece, tmp, dst=0, group=group)
            for r in range(world):
                out[r, b:e].copy_(tmp[r])
        else:
            dist.gather(piece, dst_list, dst=0, group=group)
    if rank != 0:
        return None
    return out.reshape((world * n,) + tuple(t.shape[1:]))


def gather_ragged_to_rank0(t, group=None):
    """Gather per-rank tensors whose first dimension differs (tracked boxes): sizes first, then
    padded payloads.  Returns the concatenation on rank 0, None elsewhere."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return t
    rank = dist.get_rank(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes) if sizes else 0
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, bufs, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat([bufs[r][: sizes[r]] for r in range(world)], dim=0)
