import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "comfyui-3d-pack_b200"))
import torch
from gs_b200 import rasterizer as R
dev = torch.device("cuda:0")
for n in (100_000, 1_000_000, 4_000_000):
    g = torch.Generator().manual_seed(0)
    d = torch.randn(n, 3, generator=g); d = d / d.norm(dim=1, keepdim=True) * (0.5 * torch.rand(n, 1, generator=g) ** (1 / 3))
    p = d.to(dev).contiguous()
    R.knn_mean_dist2(p); torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True); a.record(); out = R.knn_mean_dist2(p); b.record(); torch.cuda.synchronize()
    print(n, "grid ms", a.elapsed_time(b), "mean d2", float(out.mean()))
