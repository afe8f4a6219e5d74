"""Instant-NGP path measurement (BASELINE config 3 class): two L=16 hash grids, occupancy grid 64^3 pre-filled with
a soft sphere r=0.5, 1920x1080 rays, step 5e-3: sampling + encode x2 + MLP x2 + weights + accumulate, fwd+bwd.
Prints Mrays/s, Msamples/s and per-stage CUDA-event times.  Dev tool."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "comfyui-3d-pack_b200"))
import numpy as np, torch
import nerfacc
from kiui.gridencoder import GridEncoder
from kiui.nn import MLP, trunc_exp
from gs_b200 import camera

dev = torch.device("cuda:0")
H, W = (1080, 1920) if len(sys.argv) < 3 else (int(sys.argv[1]), int(sys.argv[2]))
L = 16
torch.manual_seed(0)
enc_d, enc_c = GridEncoder(num_levels=L).to(dev), GridEncoder(num_levels=L).to(dev)
mlp_d, mlp_c = MLP(2 * L, 1, 32, 2, bias=False).to(dev), MLP(2 * L, 3, 32, 2, bias=False).to(dev)
est = nerfacc.OccGridEstimator(roi_aabb=torch.tensor([-1.0, -1, -1, 1, 1, 1], device=dev), resolution=64, levels=1).to(dev)
g = (torch.arange(64, device=dev).float() + 0.5) / 64 * 2 - 1
x, y, z = torch.meshgrid(g, g, g, indexing="ij")
est.binaries = ((x * x + y * y + z * z) < 0.25)[None]; est.occs = est.binaries.flatten().float()
pose = torch.from_numpy(camera.orbit_camera(0, 30, 1.75)).to(dev)
fovy = 49.1
xs_, ys_ = torch.meshgrid(torch.arange(W, device=dev), torch.arange(H, device=dev), indexing="xy")
focal = H * 0.5 / np.tan(0.5 * np.deg2rad(fovy))
dirs = torch.nn.functional.pad(torch.stack([(xs_.flatten() - W * 0.5 + 0.5) / focal, (ys_.flatten() - H * 0.5 + 0.5) / focal * -1.0], -1), (0, 1), value=-1.0)
rd = dirs @ pose[:3, :3].T; rd = rd / rd.norm(dim=-1, keepdim=True); ro = pose[:3, 3][None].expand_as(rd).contiguous()
gimg = torch.rand(H * W, 3, device=dev)
params = list(enc_d.parameters()) + list(enc_c.parameters()) + list(mlp_d.parameters()) + list(mlp_c.parameters())

def density(p): return trunc_exp(mlp_d(enc_d(p))).squeeze(-1)
def sigma_fn(t0, t1, ri): return density(ro[ri] + rd[ri] * (t0 + t1)[:, None] / 2.0)

def step(tm=None):
    def T(name, fn):
        if tm is None: return fn()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True); a.record(); r = fn(); b.record(); tm.append((name, a, b)); return r
    def samp():
        with torch.no_grad():
            return est.sampling(ro, rd, sigma_fn=sigma_fn, near_plane=0.01, far_plane=100, render_step_size=5e-3, stratified=True, cone_angle=0)
    ri, t0, t1 = T("sampling(+density pass)", samp)
    p = ro[ri] + rd[ri] * (t0 + t1)[:, None] / 2.0
    fd = T("encode_density", lambda: enc_d(p)); sig = T("mlp_density", lambda: trunc_exp(mlp_d(fd)).squeeze(-1))
    fc = T("encode_color", lambda: enc_c(p)); rgb = T("mlp_color", lambda: torch.sigmoid(mlp_c(fc)))
    w, _, _ = T("weights", lambda: nerfacc.render_weight_from_density(t0, t1, sig, ray_indices=ri, n_rays=H * W))
    col = T("accumulate", lambda: nerfacc.accumulate_along_rays(w, values=rgb, ray_indices=ri, n_rays=H * W))
    alp = nerfacc.accumulate_along_rays(w, values=None, ray_indices=ri, n_rays=H * W)
    loss = ((col + (1 - alp)) * gimg).sum()
    T("backward", lambda: loss.backward())
    T("tv", lambda: enc_d.grad_total_variation(1e-8))
    for q in params: q.grad = None
    return ri.numel()

for _ in range(3): S = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
K = 5
e0.record()
for _ in range(K): S = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
tm = []; step(tm); torch.cuda.synchronize()
print(json.dumps({"workload": f"Instant-NGP fwd+bwd: L={L} x2 grids, {W}x{H} rays, occ 64^3 sphere r=0.5, dt 5e-3", "ms_per_step": ms,
                  "samples": S, "Mrays_per_s": H * W / ms / 1e3, "Msamples_per_s": S / ms / 1e3,
                  "op_ms": {n: a.elapsed_time(b) for n, a, b in tm}}))
