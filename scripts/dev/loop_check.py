import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "comfyui-3d-pack_b200"))
import numpy as np, torch
from gs_b200 import trainer
G = np.load(os.path.join(ROOT, "tests", "golden", "ref_loop.npz"))
dev = torch.device("cuda:0")
N, deg, K = int(G["N"]), int(G["deg"]), int(G["K"]); H, W = int(G["HW"][0]), int(G["HW"][1])
tr = trainer.GaussianTrainer(trainer.TrainParams(num_pts=N, sh_degree=deg, density_start_iter=10 ** 9), device=dev, seed=0)
t = lambda k: torch.from_numpy(G[k]).to(dev)
tr.v["xyz"].copy_(t("init_xyz")); tr.v["shs"].copy_(torch.cat([t("init_f_dc"), t("init_f_rest")], 1)); tr.v["opacity"].copy_(t("init_opacity"))
tr.v["scaling"].copy_(t("init_scaling")); tr.v["rotation"].copy_(t("init_rotation"))
tr.m1.zero_(); tr.m2.zero_(); tr.step_count = 0
ref_imgs, ref_masks = t("ref_imgs"), t("ref_masks")
lrs = {"xyz": 1.6e-3, "shs": 0.0025, "opacity": 0.05, "scaling": 0.005, "rotation": 0.001}
for s in range(K):
    i = int(G["idx"][s])
    rec = np.zeros((1, 40), dtype=np.float32)
    rec[0, :16] = G["step_view"][s].reshape(-1); rec[0, 16:32] = G["step_proj"][s].reshape(-1); rec[0, 32:35] = G["step_campos"][s]
    rec[0, 35:38] = G["step_bg"][s]; rec[0, 38:40] = G["step_tan"][s]
    loss = tr.train_step(rec, W, H, ref_imgs[i:i + 1].contiguous(), ref_masks[i:i + 1].contiguous())
    ref = {"xyz": G["after_xyz"][s], "shs": np.concatenate([G["after_f_dc"][s], G["after_f_rest"][s]], 1), "opacity": G["after_opacity"][s],
           "scaling": G["after_scaling"][s], "rotation": G["after_rotation"][s]}
    print("step", s, "loss", loss)
    for k in ref:
        d = np.abs(tr.v[k].cpu().numpy() - ref[k])
        print(f"   {k:9s} max {d.max():.3e}  p99 {np.quantile(d, 0.99):.3e}  median {np.median(d):.3e}  frac>0.1lr {float((d > 0.1 * lrs[k]).mean()):.4f}  (lr {lrs[k]})")
