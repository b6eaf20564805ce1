"""Frame sharding across the GPUs of one box (one process per GPU, torch.distributed).

The hot path is embarrassingly parallel -- every output frame depends on one source frame and one
3x3 -- so ranks share nothing and the only collective is the gather of BEV outputs / boxes to
rank 0 (NCCL over NVLink; gloo in the CPU tests).  SURVEY.md 8e.

Rank 0's NVLink ingress (900 GB/s nominal, ~770 GB/s measured for a peer copy) is what bounds the
gather: a B200 produces BEVs faster than one GPU can receive them from seven peers, so the gather
is reported separately from the warp scaling and, when it is wanted, overlapped with the warp
chunk by chunk (`pipelined_gather_to_rank0`).
"""
import contextlib

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


def _world(group):
    return dist.get_world_size(group) if dist.is_initialized() else 1


def _dst0(group):
    """Global rank of the group's rank 0 (dist.gather takes a GLOBAL destination rank)."""
    return dist.get_global_rank(group, 0) if group is not None else 0


def _first_dims(t, group):
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(_world(group))]
    dist.all_gather(sizes, n, group=group)
    return [int(s.item()) for s in sizes]


def gather_to_rank0(t, group=None, chunks=4):
    """Gather per-rank tensors on rank 0 (others get None), first dim concatenated in rank order.
    Shards that differ in their first dimension (shard_range leaves a remainder) take the padded
    route of gather_ragged_to_rank0; equal shards are moved in `chunks` pieces straight into the
    result."""
    world = _world(group)
    if world == 1:
        return t
    rank = dist.get_rank(group)
    t = t.contiguous()
    sizes = _first_dims(t, group)
    if len(set(sizes)) != 1:
        return _gather_padded(t, sizes, group)
    out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device) if rank == 0 else None
    n = t.shape[0]
    chunks = max(1, min(chunks, n))
    for c in range(chunks):
        b, e = shard_range(n, c, chunks)
        if e <= b:
            continue
        dst_list = [out[r, b:e] for r in range(world)] if rank == 0 else None
        dist.gather(t[b:e], dst_list, dst=_dst0(group), group=group)
    if rank != 0:
        return None
    return out.reshape((world * n,) + tuple(t.shape[1:]))


def _gather_padded(t, sizes, group):
    world, rank = _world(group), dist.get_rank(group)
    m = max(sizes) if sizes else 0
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, bufs, dst=_dst0(group), group=group)
    if rank != 0:
        return None
    return torch.cat([bufs[r][: sizes[r]] for r in range(world)], dim=0)


def gather_ragged_to_rank0(t, group=None):
    """Gather per-rank tensors whose first dimension differs (tracked boxes): sizes first, then
    padded payloads.  Returns the concatenation on rank 0, None elsewhere."""
    if _world(group) == 1:
        return t
    t = t.contiguous()
    return _gather_padded(t, _first_dims(t, group), group)


def pipelined_gather_to_rank0(produce, n, tail_shape, dtype, device, chunks=8, group=None, local_out=None):
    """Produce this rank's n result frames in `chunks` slices and gather every slice on rank 0
    while the next one is being produced (SURVEY.md 8e caveat 2).

    produce(b, e, out) must enqueue, on the CURRENT stream, the work that fills out[b:e] (the
    batched warp of frames b..e).  On CUDA the gather of slice k runs on a side stream that waits
    for slice k's event, so it overlaps the warp of slice k + 1; on CPU (gloo tests) the same calls
    run in order.  All ranks must hold the same n.  Returns (local_out, gathered): gathered is the
    (world * n, ...) tensor on rank 0 and None elsewhere.
    """
    world = _world(group)
    rank = dist.get_rank(group) if world > 1 else 0
    cuda = torch.device(device).type == "cuda"
    if local_out is None:
        local_out = torch.empty((n,) + tuple(tail_shape), dtype=dtype, device=device)
    gathered = None
    if world > 1 and rank == 0:
        gathered = torch.empty((world, n) + tuple(tail_shape), dtype=dtype, device=device)
    side = torch.cuda.Stream(device=device) if (cuda and world > 1) else None
    works = []
    chunks = max(1, min(chunks, n)) if n else 1
    for c in range(chunks):
        b, e = shard_range(n, c, chunks)
        if e <= b:
            continue
        produce(b, e, local_out)
        if world == 1:
            continue
        dst_list = [gathered[r, b:e] for r in range(world)] if rank == 0 else None
        if side is not None:
            ev = torch.cuda.Event()
            ev.record()
            ctx = torch.cuda.stream(side)
        else:
            ev, ctx = None, contextlib.nullcontext()
        with ctx:
            if ev is not None:
                side.wait_event(ev)
            works.append(dist.gather(local_out[b:e], dst_list, dst=_dst0(group), group=group,
                                     async_op=True))
    for w in works:
        w.wait()  # CUDA: the current stream waits for the collective; CPU: blocks
    if side is not None:
        torch.cuda.current_stream(device).wait_stream(side)
    if world == 1:
        return local_out, local_out
    return local_out, (gathered.reshape((world * n,) + tuple(tail_shape)) if rank == 0 else None)
