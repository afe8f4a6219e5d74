#!/usr/bin/env python
"""A/B of the tile-composite kernels at config-1 scale: round-1 (gs_render.cu) vs round-2 (gs_composite.cu).

Checks (same process, same inputs): images of both kernel sets bit-identical, packed gradients equal within the
parity tolerance; then times the 8-view step and the per-stage profile for both.  Prints one JSON line per section."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "comfyui-3d-pack_b200"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from gs_b200 import _lib, camera, optim_step, synthetic  # noqa: E402
from gs_b200 import rasterizer as R  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gaussians", type=int, default=1_000_000)
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--cloud", default="D0")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--skip-r1", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    N, V, W, H, deg = a.gaussians, a.views, a.width, a.height, 3
    cloud = synthetic.make_cloud(a.cloud, N, deg, seed=0, device=dev)
    params = optim_step.PackedParams(cloud)
    views_np = camera.orbit_views(V, W, H, n_total=V, start=0)
    views = optim_step.ViewSet(views_np, W, H, deg, dev)
    g = torch.Generator(device="cpu").manual_seed(1234)
    dl = torch.rand(V, 5, H, W, generator=g) * 2 - 1
    dl[:, 3:] *= 0.1
    dl = dl.to(dev)
    imgs = torch.empty(V, 5, H, W, device=dev)

    res = {}
    modes = [("r2", 0)] if a.skip_r1 else [("r1", 1), ("r2", 0)]
    for name, mode in modes:
        _lib.check(_lib.lib.gs_b200_debug_set_composite(mode))
        pairs = optim_step.step_device_pipelined(params, views, dl, imgs)
        torch.cuda.synchronize()
        res[name] = (imgs.clone(), params.grads.clone(), pairs)
        # timing
        for _ in range(3):
            optim_step.step_device_pipelined(params, views, dl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            optim_step.step_device_pipelined(params, views, dl)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        # per-stage profile through the per-view entries (culling as in the step)
        mode0 = R.get_tile_culling()
        R.set_tile_culling(2 if mode0 >= 1 else 0)
        _lib.lib.gs_b200_profile_enable(1)
        for _ in range(2):
            optim_step.step_device(params, views, dl)
        torch.cuda.synchronize()
        R.set_tile_culling(mode0)
        msv = (C.c_float * _lib.NSTAGES)(); calls = (C.c_int32 * _lib.NSTAGES)()
        _lib.check(_lib.lib.gs_b200_profile_read(msv, calls))
        _lib.lib.gs_b200_profile_enable(0)
        stage = {nm: round(msv[i] / calls[i], 4) if calls[i] else 0.0 for i, nm in enumerate(_lib.STAGE_NAMES)}
        print(json.dumps({"mode": name, "ms_per_step": ms, "msplats": N * V / ms / 1e3, "pairs_per_view": pairs / V,
                          "stage_ms_per_view": stage}), flush=True)
    if "r1" in res:
        i1, g1, _ = res["r1"]; i2, g2, _ = res["r2"]
        out = {"images_bit_equal": bool(torch.equal(i1, i2)), "image_max_abs_diff": float((i1 - i2).abs().max()),
               "image_nan": bool(torch.isnan(i2).any()), "grad_nan": bool(torch.isnan(g2).any())}
        n = N
        sizes = [("means3D", 3 * n), ("shs", 48 * n), ("opacities", n), ("scales", 3 * n), ("rotations", 4 * n), ("means2D", 3 * n)]
        o = 0
        for nm, s in sizes:
            a1, a2 = g1[o:o + s].double(), g2[o:o + s].double(); o += s
            out["grad_rel_" + nm] = float((a1 - a2).norm() / (a1.norm() + 1e-30))
            out["grad_maxabs_" + nm] = float((a1 - a2).abs().max() / (a1.abs().max() + 1e-30))
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
