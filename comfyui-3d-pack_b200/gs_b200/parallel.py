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


# ---------------------------------------------------------------------------------------------------------
# all-reduce overlapped with the tail of the step (gs_b200_set_grad_sink)
# ---------------------------------------------------------------------------------------------------------
GROUP_WIDTHS = lambda M: (3, 3 * M, 1, 3, 4, 3)          # floats per Gaussian: means3D | shs | opacities | scales | rotations | means2D


def chunk_slices(N: int, M: int, first: int, count: int):
    """(offset, length) of rows first..first+count of each of the six groups inside the packed gradient buffer."""
    out, base = [], 0
    for w in GROUP_WIDTHS(M):
        out.append((base + first * w, count * w))
        base += N * w
    return out


class OverlappedGradAllReduce:
    """One logical all-reduce of the packed gradient buffer per step, issued in Gaussian-range chunks from inside the
    step: the library calls back after each range of its last pass has been enqueued, and the range's rows of the SH
    gradient block (77 % of the buffer) are all-reduced on NCCL's stream (ordered after that kernel) while the next range
    is computed; the five narrow groups follow as one grouped collective after the last range.

        ar = OverlappedGradAllReduce(params.grads, N, M, nchunks=8)
        with ar:                       # registers / removes the sink for this thread
            step_device_pipelined(...)
        ar.wait()                      # current stream waits for every chunk

    Without an initialised process group (or world 1) it does nothing."""

    def __init__(self, grads: torch.Tensor, N: int, M: int, nchunks: int = 8, group=None, enabled: bool = True):
        self.grads, self.N, self.M, self.nchunks, self.group = grads, N, M, nchunks, group
        self.works, self.err = [], None
        # `enabled=False`: a caller that runs a single-process step while a process group exists (world passed as 1)
        self.active = bool(enabled) and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        from . import _lib
        self._lib = _lib
        self._cb = _lib.GRAD_SINK(self._sink)      # keep the ctypes thunk alive

    def _sink(self, _user, first, count, stream_ptr):
        try:
            first, count = int(first), int(count)
            with torch.cuda.stream(torch.cuda.ExternalStream(int(stream_ptr), device=self.grads.device)):
                sl = chunk_slices(self.N, self.M, first, count)
                # the SH block is 3M of the 14+3M floats per Gaussian (77 % at SH degree 3): its rows go out range by
                # range, one plain all-reduce per range; the five narrow groups follow once, whole, as one grouped launch
                # after the last range (few large collectives instead of many small ones)
                o, n = sl[1]
                self.works.append(dist.all_reduce(self.grads[o:o + n], group=self.group, async_op=True))
                if first + count >= self.N:
                    rest = [self.grads[o:o + n] for k, (o, n) in enumerate(chunk_slices(self.N, self.M, 0, self.N)) if k != 1]
                    if hasattr(dist, "_coalescing_manager"):
                        with dist._coalescing_manager(group=self.group, async_ops=True) as cm:
                            for t in rest:
                                dist.all_reduce(t, group=self.group)
                        self.works.append(cm)
                    else:
                        for t in rest:
                            self.works.append(dist.all_reduce(t, group=self.group, async_op=True))
        except BaseException as e:      # never let an exception cross the C frame
            self.err = e

    def __enter__(self):
        self.works, self.err = [], None
        if self.active:
            self._lib.check(self._lib.lib.gs_b200_set_grad_sink(self._cb, None, self.nchunks))
        return self

    def __exit__(self, *exc):
        if self.active:
            self._lib.lib.gs_b200_set_grad_sink(self._lib.GRAD_SINK(), None, 1)
        if self.err is not None and exc[0] is None:
            raise self.err
        return False

    def wait(self):
        for w in self.works:
            w.wait()
        self.works = []
