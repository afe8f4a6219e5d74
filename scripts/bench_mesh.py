"""Mesh-path measurement (BASELINE config 4 class): sphere mesh, 8 views at 1920x1080,
rasterize -> antialias(alpha) -> interpolate(uv, db) -> texture -> antialias(colour), forward + backward.
Prints Mtri*views/s, Mpixels/s and per-op CUDA-event times.  Dev tool (not the driver's bench)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "comfyui-3d-pack_b200"))
import numpy as np, torch
import nvdiffrast.torch as dr
from oracle import dr_oracle as D, gs_oracle as O      # mesh/camera generators only

dev = torch.device("cuda:0")
sub = int(sys.argv[1]) if len(sys.argv) > 1 else 7
H, W, Vn = 1080, 1920, 8
v, f, uv = D.icosphere(sub)
proj = D.gl_perspective(49.1, W / H)
pos = torch.cat([D.clip_positions(v, O.orbit_camera(0, 45.0 * k, 1.75), proj) for k in range(Vn)], dim=0).to(dev)
f = f.to(dev); uv = uv.to(dev)
pos.requires_grad_(True)
tex = torch.rand(1, 1024, 1024, 3, device=dev, requires_grad=True)
gi = torch.rand(Vn, H, W, 3, device=dev)
ctx = dr.RasterizeCudaContext()

def step(timers=None):
    def T(name, fn):
        if timers is None: return fn()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True); a.record(); r = fn(); b.record(); timers.append((name, a, b)); return r
    rast, db = T("rasterize", lambda: dr.rasterize(ctx, pos, f, (H, W)))
    alpha = T("antialias_alpha", lambda: dr.antialias(torch.clamp(rast[..., -1:], 0, 1).contiguous(), rast, pos, f))
    texc, texc_db = T("interpolate", lambda: dr.interpolate(uv[None], rast, f, rast_db=db, diff_attrs="all"))
    alb = T("texture", lambda: dr.texture(tex, texc, uv_da=texc_db, filter_mode="linear"))
    alb = T("antialias_color", lambda: dr.antialias(alb, rast, pos, f))
    loss = ((alpha * alb) * gi).sum()
    T("backward", lambda: loss.backward())
    pos.grad = None; tex.grad = None

for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
K = 5
e0.record()
for _ in range(K): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
tm = []; step(tm); torch.cuda.synchronize()
F = f.shape[0]
print(json.dumps({"workload": f"mesh fwd+bwd: {F} triangles, {Vn} views {W}x{H}", "ms_per_step": ms,
                  "Mtri_views_per_s": F * Vn / ms / 1e3, "Mpixels_per_s": Vn * H * W / ms / 1e3,
                  "op_ms": {n: a.elapsed_time(b) for n, a, b in tm}}))
