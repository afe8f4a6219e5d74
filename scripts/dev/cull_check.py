import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "comfyui-3d-pack_b200"))
import numpy as np, torch
from gs_b200 import rasterizer as R, camera
dev = torch.device("cuda:0")
N, deg, W, H = 400, 0, 320, 200
g = torch.Generator().manual_seed(3)
cloud = {"means3D": (torch.rand(N, 3, generator=g) - 0.5), "shs": torch.rand(N, 1, 3, generator=g),
         "opacities": torch.rand(N, 1, generator=g) ** 3,
         "scales": torch.exp(torch.rand(N, 3, generator=g) * 4 - 5), "rotations": torch.randn(N, 4, generator=g)}
cloud = {k: v.to(dev).contiguous() for k, v in cloud.items()}
vnp = camera.orbit_views(3, W, H)
t = lambda a: torch.from_numpy(a).to(dev)
gen = torch.Generator().manual_seed(5)
up = [torch.rand(3, H, W, generator=gen).to(dev) * 2 - 1, torch.rand(1, H, W, generator=gen).to(dev) * 0.1, torch.rand(1, H, W, generator=gen).to(dev) * 0.1]
NAMES = ("means3D", "shs", "opacities", "scales", "rotations")
for v in range(3):
    rs = R.GaussianRasterizationSettings(H, W, float(vnp[v, 38]), float(vnp[v, 39]), t(vnp[v, 35:38].copy()), 1.0,
                                        t(vnp[v, :16].copy()).view(4, 4), t(vnp[v, 16:32].copy()).view(4, 4), deg,
                                        t(vnp[v, 32:35].copy()), False, False)
    res = []
    for mode in (0, 0, 2):
        R.set_tile_culling(mode)
        leaves = {k: x.clone().requires_grad_(True) for k, x in cloud.items()}
        m2d = torch.zeros(N, 3, device=dev, requires_grad=True)
        color, radii, depth, alpha = R.GaussianRasterizer(rs)(means3D=leaves["means3D"], means2D=m2d, shs=leaves["shs"], opacities=leaves["opacities"],
                                                             scales=leaves["scales"], rotations=leaves["rotations"])
        ((color * up[0]).sum() + (depth * up[1]).sum() + (alpha * up[2]).sum()).backward()
        fs = R.forward_with_state(rs, cloud["means3D"], cloud["opacities"], shs=cloud["shs"], scales=cloud["scales"], rotations=cloud["rotations"])
        res.append(([leaves[k].grad.clone() for k in NAMES] + [m2d.grad.clone()], fs["num_rendered"]))
    print("view", v, "pairs", res[0][1], res[2][1])
    for i, nm in enumerate(NAMES + ("means2D",)):
        a, b, c = res[0][0][i], res[1][0][i], res[2][0][i]
        d = (a - c).abs()
        j = int(d.view(N, -1).max(dim=1).values.argmax())
        print(f"  {nm:10s} norm {float(a.norm()):.4e} run-to-run {float((a-b).norm()/a.norm()):.2e} culled-vs-full {float((a-c).norm()/a.norm()):.2e} worst idx {j} a={a[j].flatten()[:4].tolist()} c={c[j].flatten()[:4].tolist()}")
    R.set_tile_culling(1)
