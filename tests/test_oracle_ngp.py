"""CPU tests of the Instant-NGP oracle: table layout, hand-computable hash-grid values, marching rule on a known
grid, volume-rendering recurrences, fp64 finite differences."""
import numpy as np
import torch

from conftest import ROOT  # noqa: F401
from oracle import gs_oracle as O
from oracle import ngp_oracle as G


def test_offsets_match_package_layout():
    off = G.grid_offsets(num_levels=12)
    assert off[0] == 0 and off[1] == 4920                 # level 0: (16+1)^3 = 4913 -> padded to a multiple of 8
    assert off[2] - off[1] == 35944                        # level 1: 33^3 = 35937 -> 35944
    assert off[-1] - off[-2] == 2 ** 19 and all(o % 8 == 0 for o in off)
    assert (np.diff(off) <= 2 ** 19).all()


def test_dense_level_is_trilinear_interpolation_of_its_corners():
    off = G.grid_offsets(num_levels=1)
    emb = torch.zeros(int(off[-1]), 2)
    # put f(i,j,k) = i + 10 j + 100 k at the dense level-0 corners (res 16 -> 17^3 lattice, index i + 17 j + 289 k)
    i, j, k = torch.meshgrid(torch.arange(17), torch.arange(17), torch.arange(17), indexing="ij")
    emb[(i + 17 * j + 289 * k).reshape(-1), 0] = (i + 10 * j + 100 * k).reshape(-1).float()
    x01 = torch.tensor([[0.3, 0.6, 0.9], [0.0, 0.0, 0.0], [1.0, 1.0, 1.0]])
    out = G.grid_encode(x01 * 2 - 1, emb, off, num_levels=1)
    pos = x01 * 15 + 0.5                                    # scale = 16*2^0 - 1
    assert torch.allclose(out[:, 0], pos[:, 0] + 10 * pos[:, 1] + 100 * pos[:, 2], atol=1e-3)


def test_hashed_level_indices_follow_the_xor_prime_rule():
    off = G.grid_offsets(num_levels=8)
    x01 = torch.tensor([[0.123, 0.456, 0.789]])
    idx, w, frac = G.grid_corner_indices(x01, 7, off)       # level 7: res 2048 -> hashed
    pos = x01 * (16 * 2 ** 7 - 1) + 0.5
    g = torch.floor(pos).to(torch.int64)[0]
    expect = ((int(g[0]) * 1) ^ ((int(g[1]) * 2654435761) & 0xFFFFFFFF) ^ ((int(g[2]) * 805459861) & 0xFFFFFFFF)) % (2 ** 19)
    assert int(idx[0, 0]) == expect
    assert abs(float(w.sum()) - 1.0) < 1e-6


def test_marching_rule_on_a_half_filled_grid():
    R = 8
    binary = torch.zeros(R, R, R, dtype=torch.bool); binary[:, :, R // 2:] = True          # z >= 0 occupied
    aabb = torch.tensor([-1.0, -1, -1, 1, 1, 1])
    ro = torch.tensor([[0.05, 0.05, 3.0]]); rd = torch.tensor([[0.0, 0.0, -1.0]])         # straight down the z axis
    ri, ts, te = G.march(ro, rd, binary, aabb, near=0.01, far=100.0, dt=0.1)
    assert ri.numel() == 10                                   # enters at t=2 (z=1), leaves the occupied half at t=3 (z=0)
    assert abs(float(ts[0]) - 2.0) < 1e-6 and torch.allclose(te - ts, torch.full_like(ts, 0.1), atol=1e-6)
    assert float(te[-1]) <= 3.0 + 1e-5
    # a ray that misses the box produces nothing
    ri2, _, _ = G.march(torch.tensor([[5.0, 5.0, 3.0]]), rd, binary, aabb, 0.01, 100.0, 0.1)
    assert ri2.numel() == 0
    # stratified offset shifts every sample by the same amount
    ri3, ts3, _ = G.march(ro, rd, binary, aabb, 0.01, 100.0, 0.1, t_offset=torch.tensor([0.03]))
    assert abs(float(ts3[0]) - 2.03) < 1e-6


def test_weights_and_accumulate_recurrences():
    ts = torch.tensor([0.0, 1.0, 2.0, 0.0, 1.0]); te = ts + 1.0
    sig = torch.tensor([0.5, 1.0, 2.0, 0.1, 3.0])
    ri = torch.tensor([0, 0, 0, 2, 2])
    w, T, a = G.render_weight_from_density(ts, te, sig, ri, 3)
    a_ref = 1 - torch.exp(-sig)
    assert torch.allclose(a, a_ref)
    assert torch.allclose(T, torch.tensor([1.0, float(1 - a_ref[0]), float((1 - a_ref[0]) * (1 - a_ref[1])), 1.0, float(1 - a_ref[3])]))
    acc = G.accumulate_along_rays(w, None, ri, 3)
    assert abs(float(acc[0, 0]) - float(1 - torch.exp(-sig[:3].sum()))) < 1e-6 and float(acc[1, 0]) == 0.0
    col = G.accumulate_along_rays(w, torch.ones(5, 3), ri, 3)
    assert torch.allclose(col, acc.expand(3, 3))
    vis = G.visibility_mask(ts, te, torch.tensor([20.0, 1, 1, 0.1, 3]), ri, 3, early_stop_eps=1e-4)
    assert vis.tolist() == [True, False, False, True, True]


def test_gradients_fp64_finite_differences():
    torch.manual_seed(0)
    off = G.grid_offsets(num_levels=4)
    emb = (torch.rand(int(off[-1]), 2, dtype=torch.float64) - 0.5)
    x = torch.rand(20, 3) * 2 - 1
    gout = torch.rand(20, 8, dtype=torch.float64)
    e = emb.clone().requires_grad_(True)
    (G.grid_encode(x, e, off, num_levels=4) * gout).sum().backward()
    rng = np.random.RandomState(0)
    nz = torch.nonzero(e.grad.abs().sum(1) > 0)[:, 0]
    for j in rng.choice(nz.numpy(), size=6, replace=False):
        p = emb.clone(); m = emb.clone(); p[j, 0] += 1e-6; m[j, 0] -= 1e-6
        fd = float(((G.grid_encode(x, p, off, num_levels=4) - G.grid_encode(x, m, off, num_levels=4)) * gout).sum()) / 2e-6
        assert abs(fd - float(e.grad[j, 0])) < 1e-6 * max(1, abs(fd))
    ts = torch.tensor([0.0, 1.0, 2.0, 0.0]).double(); te = ts + 0.7
    sig = torch.tensor([0.5, 1.0, 2.0, 0.3], dtype=torch.float64, requires_grad=True)
    ri = torch.tensor([0, 0, 0, 1])
    vals = torch.rand(4, 3, dtype=torch.float64)
    def f(s):
        w, T, a = G.render_weight_from_density(ts, te, s, ri, 2)
        return (G.accumulate_along_rays(w, vals, ri, 2) * torch.tensor([[1.0, 2, 3], [4, 5, 6]])).sum() + (T * 0.3).sum() + (a * 0.2).sum()
    f(sig).backward()
    for j in range(4):
        p = sig.detach().clone(); m = sig.detach().clone(); p[j] += 1e-6; m[j] -= 1e-6
        fd = (float(f(p)) - float(f(m))) / 2e-6
        assert abs(fd - float(sig.grad[j])) < 1e-6 * max(1, abs(fd))


def test_get_rays_matches_reference_formula():
    pose = O.orbit_camera(10, 20, 2.0)
    ro, rd = G.get_rays(pose, 4, 6, 49.1)
    assert ro.shape == (24, 3) and torch.allclose(rd.norm(dim=1), torch.ones(24), atol=1e-6)
    assert torch.allclose(ro[0], torch.from_numpy(pose[:3, 3]))
    centre = rd.reshape(4, 6, 3)[1:3, 2:4].mean(dim=(0, 1))
    fwd = -torch.from_numpy(pose[:3, 2])                     # OpenGL: camera looks down -z
    assert float((centre / centre.norm() * fwd).sum()) > 0.999
