#!/bin/bash
# compute-sanitizer over the kernels added / rewritten in round 2: composite forward / backward (cp.async + mbarrier ring,
# both gathers), geometry-only preprocess + late colour kernel (host-buffer step), densification kernels, mip texture.
# Run under gpurun; writes gpurun_out/sanitizer_r2.txt.
set -o pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/sanitizer_r2.txt
cat > /tmp/r2_case.py <<'PY'
import sys, os
ROOT = os.getcwd(); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "comfyui-3d-pack_b200"))
import numpy as np, torch
from gs_b200 import camera, optim_step, synthetic, trainer
dev = torch.device("cuda:0")
N, V, W, H, deg = 6000, 3, 176, 112, 2
cloud = synthetic.make_cloud("D1", N, deg, seed=0, device=dev)
params = optim_step.PackedParams(cloud)
vnp = camera.orbit_views(V, W, H)
views = optim_step.ViewSet(vnp, W, H, deg, dev)
dl_cpu = torch.rand(V, 5, H, W, generator=torch.Generator().manual_seed(1)) * 2 - 1
img = torch.empty(V, 5, H, W, device=dev)
optim_step.step_device_pipelined(params, views, dl_cpu.to(dev), img); torch.cuda.synchronize()
print("device step grads", float(params.grads.abs().sum()), "image", float(img.sum()))
hs = optim_step.HostStep({k: v.cpu() for k, v in cloud.items()}, vnp, W, H, deg, dl_cpu)
ih = torch.empty(V, 5, H, W).pin_memory()
hs.run(images=ih); print("host step grads", float(hs.grads.abs().sum()), "images equal", bool(torch.equal(ih, img.cpu())))
if os.environ.get("R2_CASE_TRAINER", "1") == "1":
    tr = trainer.GaussianTrainer(trainer.TrainParams(num_pts=5000, sh_degree=1, density_start_iter=1, densification_interval=1,
                                                     densify_grad_threshold=1e-7, opacity_reset_interval=10 ** 9), device=dev, seed=1)
    ref = torch.rand(V, 3, 176, 176, device=dev); mask = (torch.rand(V, 1, 176, 176, device=dev) > 0.5).float()
    tv = camera.orbit_views(V, 176, 176)
    for _ in range(3):
        tr.train_step(tv, 176, 176, ref, mask)
    print("trainer N after densification", tr.N)
torch.cuda.synchronize()
PY
{
echo "# compute-sanitizer, round-2 kernels ($(date -u +%F)), scripts/sanitize_r2.sh"
echo "== memcheck (cp.async gather, default)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 python /tmp/r2_case.py 2>&1 | tail -8; echo "memcheck rc=$?"
echo "== memcheck (GS_B200_GATHER=tma)"
GS_B200_GATHER=tma R2_CASE_TRAINER=0 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 python /tmp/r2_case.py 2>&1 | tail -6; echo "memcheck tma rc=$?"
echo "== racecheck (cp.async gather, default)"
R2_CASE_TRAINER=0 timeout 900 compute-sanitizer --tool racecheck --print-limit 100000 --error-exitcode 9 --launch-timeout 0 python /tmp/r2_case.py > gpurun_out/racecheck_r2_full.txt 2>&1; echo "racecheck rc=$?"
# classify every reported hazard by the pair of code locations involved
grep -E "Race reported|and (Read|Write) access" gpurun_out/racecheck_r2_full.txt | sed -E 's/\+0x[0-9a-f]+//; s/ \[[0-9]+ hazards\]//; s/^=+ +//' | sort | uniq -c | sort -rn | head -30
grep -E "RACECHECK SUMMARY|^(device|host) step" gpurun_out/racecheck_r2_full.txt
echo "== synccheck"
R2_CASE_TRAINER=0 timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 --launch-timeout 0 python /tmp/r2_case.py 2>&1 | tail -6; echo "synccheck rc=$?"
} > $OUT 2>&1
tail -45 $OUT
