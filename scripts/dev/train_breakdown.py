import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "comfyui-3d-pack_b200"))
import numpy as np, torch
from gs_b200 import camera, trainer, optim_step
dev = torch.device("cuda:0")
N, V, W, H = 1_000_000, 8, 1920, 1080
def timeit(fn, k=5, w=2):
    for _ in range(w): fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(True), torch.cuda.Event(True); a.record()
    for _ in range(k): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / k
views = camera.orbit_views(V, W, H)
tr = trainer.GaussianTrainer(trainer.TrainParams(num_pts=N, sh_degree=3, density_start_iter=10 ** 9), device=dev, seed=0)
ref = torch.rand(V, 3, H, W, device=dev); mask = (torch.rand(V, 1, H, W, device=dev) > 0.5).float()
out = {}
img = tr.render_views(views, W, H)
out["loss_one_view_ms"] = timeit(lambda: optim_step.image_loss(img[0], ref[0], mask[0], 0.2, 3.0, 1.0))
out["loss_one_view_noss_ms"] = timeit(lambda: optim_step.image_loss(img[0], ref[0], mask[0], 0.0, 3.0, 1.0))
out["train_step_ms"] = timeit(lambda: tr.train_step(views, W, H, ref, mask))
tr.p.lambda_ssim = 0.0
out["train_step_no_ssim_ms"] = timeit(lambda: tr.train_step(views, W, H, ref, mask))
tr.p.lambda_ssim = 0.2
dl = torch.rand(V, 5, H, W, device=dev)
cl, vs = tr._cloud(), tr._viewset(views, W, H)
out["raster_step_ms"] = timeit(lambda: optim_step.step_device_pipelined(cl, vs, dl))
imgs = torch.empty(V, 5, H, W, device=dev); lv = torch.zeros(V, device=dev); rad = torch.empty(V, N, dtype=torch.int32, device=dev)
out["raster_step_train_entry_ms"] = timeit(lambda: optim_step.step_device_train(cl, vs, ref, mask, 0.2, 3.0, 1.0 / V, imgs, dl, lv, rad))
out["raster_step_keep_images_ms"] = timeit(lambda: optim_step.step_device_pipelined(cl, vs, dl, imgs))
out["activate_ms"] = timeit(tr.activate)
from gs_b200 import _lib
import ctypes as C
from gs_b200.rasterizer import _ptr, _stream
lrs = tr.learning_rates(0)
out["adam_ms"] = timeit(lambda: _lib.check(_lib.lib.gs_b200_adam_step(tr.N, tr.M, C.c_void_p(lrs.ctypes.data), 0.9, 0.999, 1e-15, 5, 1.0, _ptr(tr.grads), _ptr(tr.raw), _ptr(tr.m1), _ptr(tr.m2), _stream())))
print(json.dumps(out))
