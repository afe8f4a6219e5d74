"""Seeded synthetic Gaussian clouds for benchmarks.

D0 = the reference's own random initialisation (main_3DGS_renderer.py:811-826 feeding
create_from_pcd :407-433): uniform ball r=0.5, isotropic scale = sqrt(mean 3-NN d^2),
identity rotation, opacity 0.1, near-grey DC colour, higher SH zero — with the 3-NN
distances from this package's own distCUDA2 replacement (gs_b200_knn_mean_dist2).
D1 = trained-like (anisotropic log-normal scales, random rotations, spread opacities).
Returned tensors are ACTIVATED (what the rasterizer boundary receives), fp32, on `device`.
"""
import numpy as np
import torch

from .rasterizer import knn_mean_dist2

SH_C0 = 0.28209479177387814


def make_cloud(kind: str, n: int, sh_degree: int, seed: int = 0, device="cuda"):
    rng = np.random.RandomState(seed)
    phis = rng.random_sample((n,)) * 2 * np.pi
    costheta = rng.random_sample((n,)) * 2 - 1
    thetas = np.arccos(costheta)
    mu = rng.random_sample((n,))
    radius = 0.5 * np.cbrt(mu)
    xyz = np.stack((radius * np.sin(thetas) * np.cos(phis), radius * np.sin(thetas) * np.sin(phis),
                    radius * np.cos(thetas)), axis=1).astype(np.float32)
    M = (sh_degree + 1) ** 2
    shs = np.zeros((n, M, 3), dtype=np.float32)
    col = (rng.random_sample((n, 3)) / 255.0) * SH_C0 + 0.5
    shs[:, 0, :] = ((col - 0.5) / SH_C0).astype(np.float32)
    means = torch.from_numpy(xyz).to(device)
    d2 = torch.clamp_min(knn_mean_dist2(means), 1e-7)
    if kind == "D0":
        scales = torch.sqrt(d2)[:, None].repeat(1, 3).contiguous()
        rots = torch.zeros(n, 4, device=device); rots[:, 0] = 1
        opac = torch.full((n, 1), 0.1, device=device)
    elif kind == "D1":
        med = float(torch.sqrt(d2).median())
        scales = torch.from_numpy(np.exp(rng.normal(np.log(med), 0.5, size=(n, 3))).astype(np.float32)).to(device)
        q = rng.normal(size=(n, 4))
        rots = torch.from_numpy((q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)).to(device)
        opac = torch.from_numpy((1 / (1 + np.exp(-rng.normal(0, 2, size=(n, 1))))).astype(np.float32)).to(device)
        shs[:, 0, :] = rng.normal(0, 1.0, size=(n, 3)).astype(np.float32)
        if M > 1:
            shs[:, 1:, :] = rng.normal(0, 0.05, size=(n, M - 1, 3)).astype(np.float32)
    else:
        raise ValueError(kind)
    return dict(means3D=means, shs=torch.from_numpy(shs).to(device), opacities=opac, scales=scales, rotations=rots)
