"""Full optimisation step at BASELINE config-1 scale: GaussianTrainer.train_step = activation + V x (forward ->
L1 + alpha-MSE + MS-SSIM loss and its gradient -> backward) in one pipeline + fused chain-rule/Adam.  Dev tool; the
headline metric (bench.py) excludes loss and optimizer by definition and reports this separately."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "comfyui-3d-pack_b200"))
import numpy as np, torch
from gs_b200 import camera, trainer

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
V = int(sys.argv[2]) if len(sys.argv) > 2 else 8
W, H = 1920, 1080
dev = torch.device("cuda:0")
views = camera.orbit_views(V, W, H)
tr = trainer.GaussianTrainer(trainer.TrainParams(num_pts=N, sh_degree=3, density_start_iter=10 ** 9), device=dev, seed=0)
ref = torch.rand(V, 3, H, W, device=dev); mask = (torch.rand(V, 1, H, W, device=dev) > 0.5).float()
for _ in range(3):
    tr.train_step(views, W, H, ref, mask)
torch.cuda.synchronize()
K = 5
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(K):
    loss = tr.train_step(views, W, H, ref, mask)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
# render-only multi-view entry
for _ in range(2):
    tr.render_views(views, W, H)
torch.cuda.synchronize(); e0.record()
for _ in range(K):
    tr.render_views(views, W, H)
e1.record(); torch.cuda.synchronize()
ms_r = e0.elapsed_time(e1) / K
print(json.dumps({"workload": f"full train step: {N} Gaussians SH3, {V} views {W}x{H}, L1+alphaMSE+MS-SSIM, fused Adam",
                  "ms_per_step": ms, "Msplats_per_s_full_step": N * V / ms / 1e3, "loss": loss,
                  "render_only_ms": ms_r, "render_only_Msplats_per_s": N * V / ms_r / 1e3, "render_fps_1080p": V / ms_r * 1e3}))
