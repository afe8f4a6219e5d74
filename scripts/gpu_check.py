"""Quick on-GPU sanity run (dev tool): sort, forward parity, backward parity vs the oracle."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "comfyui-3d-pack_b200"))
import numpy as np, torch
from oracle import gs_oracle as O
from gs_b200 import rasterizer as R

dev = torch.device("cuda:0")
torch.manual_seed(0)
# ---- sort ----
for n, bits in [(1, 32), (1000, 32), (4096, 32), (4097, 13), (1 << 20, 32), (3_000_001, 13)]:
    k = torch.randint(0, 2 ** 31 - 1, (n,), device=dev, dtype=torch.int32)
    if bits < 32: k = k & ((1 << bits) - 1)
    v = torch.arange(n, device=dev, dtype=torch.int32)
    sk, sv = R.sort_pairs_u32(k, v, 0, bits)
    rk, ri = torch.sort(k.to(torch.int64), stable=True)
    print("sort", n, bits, "keys ok", bool((sk.to(torch.int64) == rk).all()), "vals ok", bool((sv.to(torch.int64) == ri).all()))

def run_case(kind, N, deg, W, H, az, seed):
    cl = O.make_cloud(kind, N, deg, seed=seed)
    st = O.minicam_settings(O.orbit_camera(0, az, 1.75), W, H, 49.1, sh_degree=deg)
    g = torch.Generator().manual_seed(seed)
    dc = torch.rand(3, H, W, generator=g) * 2 - 1
    dd = (torch.rand(1, H, W, generator=g) * 2 - 1) * 0.1
    da = (torch.rand(1, H, W, generator=g) * 2 - 1) * 0.1
    t = time.time()
    out, grads = O.rasterize_with_grads({k: cl[k] for k in ("means3D", "shs", "opacities", "scales", "rotations")}, st, dc, dd, da)
    t_or = time.time() - t
    rs = R.GaussianRasterizationSettings(H, W, st.tanfovx, st.tanfovy, st.bg.to(dev), 1.0, st.viewmatrix.to(dev),
                                        st.projmatrix.to(dev), deg, st.campos.to(dev), False, True)
    inp = {k: cl[k].to(dev).requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    fs = R.forward_with_state(rs, inp["means3D"].detach(), inp["opacities"].detach(), shs=inp["shs"].detach(),
                              scales=inp["scales"].detach(), rotations=inp["rotations"].detach())
    aux = out["aux"]
    print(f"[{kind} N={N} deg={deg} {W}x{H}] P={fs['num_rendered']} oracle P={aux['keys'].size} oracle_time={t_or:.2f}s")
    print("  radii exact", bool((fs["radii"].cpu() == out["radii"]).all()),
          "keys exact", fs["num_rendered"] == aux["keys"].size and bool((fs["sorted_keys"].cpu().numpy().astype(np.uint64) == aux["keys"]).all()),
          "point_list exact", bool((fs["point_list"].cpu().numpy().astype(np.uint32) == aux["point_list"]).all()),
          "ranges exact", bool((fs["ranges"].cpu().numpy().astype(np.uint32) == aux["ranges"]).all()),
          "n_contrib mismatches", int((fs["n_contrib"].cpu() != aux["n_contrib"]).sum()))
    print("  color max abs", float((fs["color"].cpu() - out["color"]).abs().max()),
          "depth", float((fs["depth"].cpu() - out["depth"]).abs().max()),
          "alpha", float((fs["alpha"].cpu() - out["alpha"]).abs().max()))
    m2 = torch.zeros(N, 3, device=dev, requires_grad=True)
    color, radii, depth, alpha = R.GaussianRasterizer(rs)(means3D=inp["means3D"], means2D=m2, shs=inp["shs"], colors_precomp=None,
                                                         opacities=inp["opacities"], scales=inp["scales"], rotations=inp["rotations"], cov3D_precomp=None)
    loss = (color * dc.to(dev)).sum() + (depth * dd.to(dev)).sum() + (alpha * da.to(dev)).sum()
    loss.backward()
    for k in ("means3D", "shs", "opacities", "scales", "rotations"):
        a, b = inp[k].grad.cpu(), grads[k]
        print(f"  grad {k:10s} max|ref| {float(b.abs().max()):.3e} max abs err {float((a-b).abs().max()):.3e} rel(norm) {float((a-b).norm()/(b.norm()+1e-30)):.3e}")
    a, b = m2.grad.cpu(), grads["means2D"]
    print(f"  grad means2D    max|ref| {float(b.abs().max()):.3e} max abs err {float((a-b).abs().max()):.3e} rel(norm) {float((a-b).norm()/(b.norm()+1e-30)):.3e}")

run_case("D0", 2000, 0, 128, 128, 0, 0)
run_case("D1", 2000, 3, 128, 128, 30, 1)
run_case("D1", 20000, 2, 200, 120, 75, 2)
run_case("D0", 50000, 3, 480, 270, 45, 0)

# ---- multi-view step entries agree with the per-view path -------------------------------------------
from gs_b200 import camera, optim_step, synthetic
N, V, W, H, deg = 30000, 5, 320, 200, 2
cloud = synthetic.make_cloud("D1", N, deg, seed=5, device=dev)
params = optim_step.PackedParams(cloud)
vnp = camera.orbit_views(V, W, H)
views = optim_step.ViewSet(vnp, W, H, deg, dev)
g = torch.Generator().manual_seed(7)
dl_cpu = torch.rand(V, 5, H, W, generator=g) * 2 - 1
dl = dl_cpu.to(dev)
imgs_a = torch.empty(V, 5, H, W, device=dev); imgs_b = torch.empty(V, 5, H, W, device=dev)
p1 = optim_step.step_device(params, views, dl, imgs_a); g1 = params.grads.clone()
for it in range(3):
    p2 = optim_step.step_device_pipelined(params, views, dl, imgs_b); g2 = params.grads.clone()
    print("pipelined step: pairs", p1, p2, "images max diff", float((imgs_a - imgs_b).abs().max()),
          "grads rel", float((g1 - g2).norm() / g1.norm()), "max abs", float((g1 - g2).abs().max()))
hs = optim_step.HostStep({k: v.cpu() for k, v in cloud.items()}, vnp, W, H, deg, dl_cpu)
for it in range(2):
    p3 = hs.run()
    print("host step: pairs", p3, "grads rel", float((hs.grads.to(dev) - g1).norm() / g1.norm()))
