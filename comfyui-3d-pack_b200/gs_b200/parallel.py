"""Data-parallel multi-view optimisation: Gaussians replicated on every rank, camera views
sharded, ONE all-reduce of the packed gradient buffer per step.  New functionality relative to the reference
(single GPU, one view at a time, main_3DGS.py:158-174); the loss terms there are batch means, so with equal
shards the global gradient is the rank-sum divided by world (average=True)."""
import torch
import torch.distributed as dist


def shard_views(n_views_total: int, rank: int, world: int):
    """Round-robin deal of the global view batch: rank r renders views r, r+world, ..."""
    return list(range(rank, n_views_total, world))


def pack_grads(grads: dict, order=("means3D", "shs", "opacities", "scales", "rotations", "means2D")) -> torch.Tensor:
    """Packed layout of gs_b200_step_device: means3D | shs | opacities | scales | rotations | means2D."""
    return torch.cat([grads[k].reshape(-1).float() for k in order])


def unpack_grads(buf: torch.Tensor, N: int, M: int) -> dict:
    sizes = [("means3D", (N, 3)), ("shs", (N, M, 3)), ("opacities", (N, 1)), ("scales", (N, 3)),
             ("rotations", (N, 4)), ("means2D", (N, 3))]
    out, o = {}, 0
    for k, shp in sizes:
        n = 1
        for s in shp:
            n *= s
        out[k] = buf[o:o + n].view(*shp); o += n
    assert o == buf.numel()
    return out


def allreduce_packed_grads(buf: torch.Tensor, average: bool = False, group=None) -> torch.Tensor:
    """One collective per step over the single contiguous buffer (NCCL on GPUs, gloo in CPU tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        if average:
            buf /= dist.get_world_size(group)
    return buf


def allreduce_densify_stats(grad_accum: torch.Tensor, denom: torch.Tensor, max_radii2D: torch.Tensor, group=None):
    """Keep replicas identical through densification (SURVEY §8e "extra state"): the per-Gaussian statistics that
    decide clone / split / prune (`xyz_gradient_accum`, `denom`, `max_radii2D`, main_3DGS_renderer.py:767-769,
    main_3DGS.py:212) are accumulated per rank over that rank's views; before each densification they are combined —
    SUM for the two accumulators (packed into one collective), MAX for the radii — so every rank takes the same
    decisions.  In place; no-op for a single process."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return
    packed = torch.cat([grad_accum.reshape(-1), denom.reshape(-1)])
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    n = grad_accum.numel()
    grad_accum.copy_(packed[:n].view_as(grad_accum)); denom.copy_(packed[n:].view_as(denom))
    dist.all_reduce(max_radii2D, op=dist.ReduceOp.MAX, group=group)
