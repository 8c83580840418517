"""Multi-GPU plumbing: one process per GPU, clouds sharded across ranks, ONE collective.

Every op of the forward pass is per-cloud (the batch index is only ever a grid dimension;
inference BatchNorm uses stored statistics), so the path shards by independent clouds with no
data-path collective.  The only exchange is the all-gather of the [B_local, 256] global descriptors
at the end (config 4 of BASELINE.json: 4096 clouds -> [4096,256] on every rank, 4 MiB), issued
through torch.distributed (NCCL on GPUs; gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous block partition of ``total`` clouds: rank r owns [lo, hi).  Blocks differ by at
    most one cloud and concatenating the ranks' blocks in rank order restores the global order."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_descriptors(local_desc, counts=None, group=None):
    """local_desc [B_local, D] on every rank -> [sum B_local, D] in rank order on every rank.
    ``counts`` (list of per-rank row counts) is needed only when the ranks hold different numbers
    of clouds; equal counts use a single all_gather_into_tensor."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_desc
    world = dist.get_world_size(group)
    local_desc = local_desc.contiguous()
    if counts is None or len(set(counts)) == 1:
        out = torch.empty((world * local_desc.shape[0],) + tuple(local_desc.shape[1:]),
                          dtype=local_desc.dtype, device=local_desc.device)
        dist.all_gather_into_tensor(out, local_desc, group=group)
        return out
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(local_desc.shape[1:]), dtype=local_desc.dtype,
                      device=local_desc.device)
    pad[:local_desc.shape[0]] = local_desc
    out = torch.empty((world * mx,) + tuple(local_desc.shape[1:]), dtype=local_desc.dtype,
                      device=local_desc.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    out = out.view(world, mx, *local_desc.shape[1:])
    return torch.cat([out[r, :counts[r]] for r in range(world)], dim=0)


def extract_global_descriptors(model, clouds, micro_batch=32, group=None):
    """Sharded global-descriptor extraction (the role of evaluate/global_eval/globaldesc_extract.py:
    61-119 in the reference, minus disk IO): ``clouds`` is the FULL [T,N,3] host tensor (every rank
    sees the same list); each rank runs its block in micro-batches and the descriptors are
    all-gathered so every rank returns the full [T, 256] matrix in the original order."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    total = clouds.shape[0]
    lo, hi = shard_range(total, rank, world)
    device = next(model.parameters()).device
    outs = []
    for s in range(lo, hi, micro_batch):
        batch = clouds[s:min(hi, s + micro_batch)].to(device, non_blocking=True)
        outs.append(model(batch, outputs=("globaldesc",))["globaldesc"])
    local = torch.cat(outs, 0) if outs else torch.empty((0, 256), device=device)
    counts = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]
    return all_gather_descriptors(local, counts, group)
