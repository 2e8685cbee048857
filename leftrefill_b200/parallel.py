"""Data-parallel plumbing for the sampler: one process per GPU, stitched canvases sharded across ranks, weights
replicated, exactly one collective per batch (an all-gather of the outputs). SURVEY §8e.

Every canvas (and its cond/uncond CFG pair) is independent for all DDIM steps — there is no cross-sample op in
UNetModel.forward or p_sample_ddim — so nothing is exchanged inside the loop. For the multiview UNet the unit is one
sample (its `group` UNet-batch rows must stay on one rank). Noise is drawn for the GLOBAL batch from one seeded CPU
generator and sliced, so results do not depend on the world size.
"""
import os

import torch
import torch.distributed as dist


def shard_bounds(total, rank, world, group=1):
    """Contiguous [lo, hi) slice of `total` items for `rank`; items come in indivisible groups of `group`."""
    assert total % group == 0, "batch must be a whole number of groups"
    units = total // group
    base, rem = divmod(units, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo * group, hi * group


def init_distributed(backend=None):
    """torchrun-style rendezvous (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT). Returns (rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world == 1:
        return 0, 1
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def global_randn(shape, seed, lo=None, hi=None, device="cpu"):
    """Noise for the global batch from one seeded CPU generator; optional [lo, hi) slice along dim 0."""
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    t = torch.randn(shape, generator=g)
    if lo is not None:
        t = t[lo:hi]
    return t.to(device)


def gather_outputs(local, total, rank, world, group=1):
    """All-gather per-rank output shards [n_local, ...] into [total, ...] on every rank (the one collective of a
    batch). Ragged shards are padded to the largest shard for the collective and trimmed afterwards."""
    if world == 1:
        return local
    sizes = [shard_bounds(total, r, world, group) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous())
    parts = [out[r * mx: r * mx + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(parts, dim=0)


def max_over_ranks(value, device):
    """Device-timed duration reported as the max over ranks."""
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sample_sharded(sampler_factory, S, global_batch, shape, cond, ucond, x_T, step_noise, rank, world, group=1,
                   **sample_kwargs):
    """Runs DDIMSampler.sample on this rank's shard of a global batch and gathers the results.

    cond / ucond: dicts {'c_concat': [T], 'c_crossattn': [T]} holding GLOBAL-batch tensors; x_T [B, C, H, W] and
    step_noise [S, B, C, H, W] (or None) are global too. Returns the gathered samples [B, C, H, W].
    """
    lo, hi = shard_bounds(global_batch, rank, world, group)
    cut = lambda c: None if c is None else {k: [t[lo:hi] for t in v] for k, v in c.items()}
    sampler = sampler_factory()
    if step_noise is not None:
        sampler.noise_source = lambda shp, dev, i: step_noise[i, lo:hi].to(dev)
    if hi > lo:
        local, _ = sampler.sample(S, hi - lo, shape, cut(cond), x_T=x_T[lo:hi], unconditional_conditioning=cut(ucond),
                                  verbose=False, **sample_kwargs)
    else:
        local = x_T[:0]
    return gather_outputs(local.contiguous(), global_batch, rank, world, group)
