"""Multi-rank GPU tests (pytest -m gpu, needs >= 2 GPUs; skipped on a single-GPU box): two NCCL ranks, views sharded as
`bench.py --gpus N` and `GaussianTrainer` shard them (replacing the per-view loop of
MVs_Algorithms/GaussianSplatting/main_3DGS.py:158-174).

 * the all-reduced packed gradient buffer — through the overlapped, chunked path (gs_b200_set_grad_sink) and through
   one plain all-reduce — equals the single-process sum over the union of the views to 1e-5 relative (pre-Adam);
 * `GaussianTrainer` replicas stay bit-identical through two densifications, the pre-Adam gradient of a 2-rank step
   equals the single-process step over the same views, and the first densification makes the same number of points.
"""
import os
import socket
import sys

import pytest
import torch

from conftest import ROOT  # noqa: F401  (sys.path setup)

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
        import numpy as np
        import torch.distributed as dist
        from gs_b200 import camera, optim_step, parallel, synthetic, trainer
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        out = {}

        # ---------------- (1) rasterizer step: all-reduced gradients == single-process sum over the union of the views
        N, W, H, deg, V = 60_000, 640, 360, 2, 6
        M = (deg + 1) ** 2
        cloud = synthetic.make_cloud("D1", N, deg, seed=11, device=dev)
        vnp = camera.orbit_views(V, W, H)
        g = torch.Generator().manual_seed(5)
        dl_all = torch.rand(V, 5, H, W, generator=g) * 2 - 1
        dl_all[:, 3:] *= 0.1
        mine = parallel.shard_views(V, rank, world)
        params = optim_step.PackedParams(cloud)
        views = optim_step.ViewSet(np.ascontiguousarray(vnp[mine]), W, H, deg, dev)
        dl = dl_all[mine].contiguous().to(dev)
        ar = parallel.OverlappedGradAllReduce(params.grads, N, M, nchunks=4)
        with ar:
            optim_step.step_device_pipelined(params, views, dl)
        ar.wait()
        torch.cuda.synchronize()
        g_overlap = params.grads.clone()
        optim_step.step_device_pipelined(params, views, dl)
        parallel.allreduce_packed_grads(params.grads)
        torch.cuda.synchronize()
        g_plain = params.grads.clone()
        both = [torch.zeros_like(g_overlap) for _ in range(world)]
        dist.all_gather(both, g_overlap)
        out["overlap_replicas_equal"] = all(torch.equal(both[0], b) for b in both)
        if rank == 0:
            vall = optim_step.ViewSet(vnp, W, H, deg, dev)
            optim_step.step_device_pipelined(params, vall, dl_all.to(dev))
            torch.cuda.synchronize()
            ref = params.grads.double()
            o = 0
            for nm, wdt in zip(("means3D", "shs", "opacities", "scales", "rotations", "means2D"), parallel.GROUP_WIDTHS(M)):
                r = ref[o:o + N * wdt]
                out["rel_overlap_" + nm] = float((g_overlap[o:o + N * wdt].double() - r).norm() / r.norm())
                out["rel_plain_" + nm] = float((g_plain[o:o + N * wdt].double() - r).norm() / r.norm())
                o += N * wdt

        # ---------------- (2) trainer: replicas identical through densification; 2-rank step == single-process step
        Wt = Ht = 176
        Vt = 4
        tv = camera.orbit_views(Vt, Wt, Ht)
        g = torch.Generator().manual_seed(1)
        ref_img = torch.rand(Vt, 3, Ht, Wt, generator=g).to(dev)
        mask = (torch.rand(Vt, 1, Ht, Wt, generator=g) > 0.4).float().to(dev)

        def make():
            tr = trainer.GaussianTrainer(trainer.TrainParams(num_pts=4000, sh_degree=1, density_start_iter=2, densification_interval=2,
                                                             densify_grad_threshold=1e-6, opacity_reset_interval=10 ** 9), device=dev, seed=3)
            tr.v["shs"][:, 0, :] = torch.rand(tr.N, 3, generator=torch.Generator().manual_seed(2)).to(dev)
            tr.v["opacity"].fill_(0.3)
            return tr
        tmine = parallel.shard_views(Vt, rank, world)
        torch.manual_seed(7)
        tr = make()
        n_hist, grads_step0 = [], None
        for s in range(5):
            tr.train_step(tv[tmine], Wt, Ht, ref_img[tmine].contiguous(), mask[tmine].contiguous())
            if s == 0:
                grads_step0 = tr.grads.clone()          # all-reduced, pre-Adam scaling (world x global mean)
            n_hist.append(tr.N)
        ns = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(ns, torch.tensor([tr.N], device=dev))
        raws = [torch.zeros_like(tr.raw) for _ in range(world)]
        dist.all_gather(raws, tr.raw)
        out["trainer_replicas_identical"] = all(int(x) == tr.N for x in ns) and all(torch.equal(raws[0], r) for r in raws)
        out["n_hist_dp"] = n_hist
        if rank == 0:
            torch.manual_seed(7)
            t1 = make()
            n1 = []
            for s in range(3):
                t1.train_step(tv, Wt, Ht, ref_img, mask, world=1)
                if s == 0:
                    g1 = t1.grads.clone()
                n1.append(t1.N)
            # DP buffer = world x (global-mean gradient); the single process holds the global-mean gradient
            a, b = grads_step0.double() / world, g1.double()
            out["trainer_grad_rel"] = float((a - b).norm() / b.norm())
            out["n_hist_single"] = n1
        dist.barrier()
        if rank == 0:
            q.put(out)
        dist.destroy_process_group()
    except BaseException as e:   # noqa: BLE001
        import traceback
        q.put({"error": "rank %d: %s\n%s" % (rank, e, traceback.format_exc())})
        raise


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl_step_and_trainer_match_the_single_process_run():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        out = q.get(timeout=420)
    except Exception:  # noqa: BLE001  (a rank hung or died: do not leave NCCL kernels spinning on the GPUs)
        for p in procs:
            p.kill()
        raise
    for p in procs:
        p.join(timeout=120)
        if p.is_alive():
            p.kill()
    assert "error" not in out, out.get("error")
    log_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(log_dir):
        import json
        with open(os.path.join(log_dir, "dp_test_result.json"), "w") as f:
            json.dump(out, f, indent=1)
    assert all(p.exitcode == 0 for p in procs)
    assert out["overlap_replicas_equal"]
    for k, v in out.items():
        if k.startswith("rel_"):
            # same terms, different fp32 summation order (atomics per view, then rank sum vs one 6-view sum); scale and
            # rotation gradients are sums of cancelling terms (test_gpu_parity uses 1e-3 for the same comparison).
            # Measured values are written to gpurun_out/dp_test_result.json.
            assert v < (1e-3 if k.endswith(("rotations", "scales")) else 1e-4), (k, v)
    assert out["trainer_replicas_identical"]
    assert out["trainer_grad_rel"] < 1e-4, out["trainer_grad_rel"]
    # the first densification (after step index 2) adds the same number of points as the single-process run
    assert out["n_hist_dp"][:3] == out["n_hist_single"][:3], (out["n_hist_dp"], out["n_hist_single"])
    assert out["n_hist_dp"][2] > out["n_hist_dp"][1]
