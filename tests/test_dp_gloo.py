"""world_size-2 gloo test (CPU) of the data-parallel host logic: view sharding + one all-reduce of the packed
gradient buffer equals the single-process sum over all views.  Per-view gradients come from the oracle."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT  # noqa: F401  (sys.path setup)
from gs_b200 import parallel
from oracle import gs_oracle as O

N, W, H, DEG, V = 300, 48, 32, 1, 4
NAMES = ("means3D", "shs", "opacities", "scales", "rotations")


def _view_grads(v):
    cl = O.make_cloud("D1", N, DEG, seed=4)
    st = O.minicam_settings(O.orbit_camera(0, 360.0 * v / V, 1.75), W, H, 49.1, sh_degree=DEG)
    g = torch.Generator().manual_seed(100 + v)
    dc = torch.rand(3, H, W, generator=g) * 2 - 1
    dd = torch.rand(1, H, W, generator=g) * 0.1
    da = torch.rand(1, H, W, generator=g) * 0.1
    _, grads = O.rasterize_with_grads({k: cl[k] for k in NAMES}, st, dc, dd, da)
    return parallel.pack_grads(grads)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = parallel.shard_views(V, rank, world)
    buf = torch.zeros(N * (3 + 3 * (DEG + 1) ** 2 + 1 + 3 + 4 + 3))
    for v in mine:
        buf += _view_grads(v)
    parallel.allreduce_packed_grads(buf)
    # densification statistics: SUM of accumulators, MAX of radii -> identical on every rank afterwards
    g = torch.Generator().manual_seed(rank)
    acc, den, rad = torch.rand(N, generator=g), torch.rand(N, generator=g).round(), torch.rand(N, generator=g) * 9
    parallel.allreduce_densify_stats(acc, den, rad)
    if rank == 0:
        q.put((mine, buf.numpy(), acc.numpy(), den.numpy(), rad.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_views_partition():
    for world in (1, 2, 3, 8):
        allv = sorted(v for r in range(world) for v in parallel.shard_views(200, r, world))
        assert allv == list(range(200))
        sizes = [len(parallel.shard_views(200, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip():
    M = (DEG + 1) ** 2
    g = {"means3D": torch.randn(N, 3), "shs": torch.randn(N, M, 3), "opacities": torch.randn(N, 1),
         "scales": torch.randn(N, 3), "rotations": torch.randn(N, 4), "means2D": torch.randn(N, 3)}
    u = parallel.unpack_grads(parallel.pack_grads(g), N, M)
    for k in g:
        assert torch.equal(u[k], g[k])


def test_two_rank_allreduce_equals_single_process_sum():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    mine, got, acc, den, rad = q.get(timeout=300)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert mine == [0, 2]
    ref = sum(_view_grads(v) for v in range(V)).numpy()
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-6)
    gens = [torch.Generator().manual_seed(r) for r in range(2)]
    parts = [(torch.rand(N, generator=g), torch.rand(N, generator=g).round(), torch.rand(N, generator=g) * 9) for g in gens]
    assert np.allclose(acc, (parts[0][0] + parts[1][0]).numpy()) and np.allclose(den, (parts[0][1] + parts[1][1]).numpy())
    assert np.array_equal(rad, torch.maximum(parts[0][2], parts[1][2]).numpy())
