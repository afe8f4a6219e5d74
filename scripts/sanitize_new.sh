#!/bin/bash
# compute-sanitizer over the kernels added late in round 1 (tile culling, in-pipeline loss, grid kNN, run-reduced
# hash-grid scatter, flat Adam); run under gpurun.
set -o pipefail
cd "$(dirname "$0")/.."
cat > /tmp/new_case.py <<'PY'
import sys, os
ROOT = os.getcwd(); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "comfyui-3d-pack_b200"))
import numpy as np, torch
from gs_b200 import camera, optim_step, synthetic, trainer, rasterizer as R
dev = torch.device("cuda:0")
W = H = 176
tr = trainer.GaussianTrainer(trainer.TrainParams(num_pts=9000, sh_degree=1, density_start_iter=10 ** 9), device=dev, seed=1)   # 9000 > 8192: grid kNN
views = camera.orbit_views(3, W, H)
ref = torch.rand(3, 3, H, W, device=dev); mask = (torch.rand(3, 1, H, W, device=dev) > 0.5).float()
print("train loss", tr.train_step(views, W, H, ref, mask))                       # culled lists, loss kernels, Adam
rad = torch.empty(3, tr.N, dtype=torch.int32, device=dev)
img = tr.render_views(views, W, H, radii=rad); print("render", float(img.sum()))
l, dl = optim_step.image_loss(img[0].contiguous(), ref[0].contiguous(), mask[0].contiguous(), 0.2, 3.0, 1.0); print("loss", float(l))
R.set_tile_culling(2)
cloud = synthetic.make_cloud("D1", 3000, 1, seed=0, device=dev)
leaves = {k: v.clone().requires_grad_(True) for k, v in cloud.items()}
t = lambda a: torch.from_numpy(a).to(dev)
rs = R.GaussianRasterizationSettings(H, W, float(views[0, 38]), float(views[0, 39]), t(views[0, 35:38].copy()), 1.0, t(views[0, :16].copy()).view(4, 4),
                                    t(views[0, 16:32].copy()).view(4, 4), 1, t(views[0, 32:35].copy()), False, False)
c, r, d, a = R.GaussianRasterizer(rs)(means3D=leaves["means3D"], means2D=torch.zeros(3000, 3, device=dev, requires_grad=True), shs=leaves["shs"],
                                      opacities=leaves["opacities"], scales=leaves["scales"], rotations=leaves["rotations"])
(c.sum() + a.sum()).backward(); print("culled single-view grads", float(leaves["means3D"].grad.abs().sum()))
R.set_tile_culling(1)
from kiui.gridencoder import GridEncoder
enc = GridEncoder(num_levels=8).to(dev)
o = torch.rand(40, 3, device=dev) - 0.5; dd_ = torch.nn.functional.normalize(torch.randn(40, 3, device=dev), dim=1)
x = (o[:, None] + dd_[:, None] * torch.linspace(0, 0.5, 100, device=dev)[None, :, None]).reshape(-1, 3).clamp(-1, 1).contiguous()
out = enc(x); out.backward(torch.rand_like(out)); print("ngp grad", float(enc.embeddings.grad.abs().sum()))
torch.cuda.synchronize()
PY
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 python /tmp/new_case.py 2>&1 | tail -12
echo "memcheck rc=$?"
