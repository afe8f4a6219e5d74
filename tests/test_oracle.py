"""CPU tests of the oracle (no GPU): pinned against the reference's golden vectors where the reference has
importable code for the path, hand-computable micro cases, sequential-vs-vectorised composite, and
autograd-vs-finite-difference in fp64."""
import math
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import gs_oracle as O


def _settings(W=64, H=64, el=0, az=0, deg=0, bg=(0., 0., 0.), fovy=49.1):
    return O.minicam_settings(O.orbit_camera(el, az, 1.75), W, H, fovy, bg=bg, sh_degree=deg)


# ---------------- golden vectors produced by the reference's own code ------------------------------------
def test_sh_basis_matches_reference_eval_sh():
    g = np.load(os.path.join(GOLDEN, "ref_sh.npz"))
    dirs = torch.from_numpy(g["dirs"])
    for deg in range(4):
        sh = torch.from_numpy(g[f"sh{deg}"])                       # reference layout [N,3,M]
        basis = O.sh_basis(deg, dirs)                              # [N,M]
        rgb = (basis[:, None, :] * sh).sum(-1)
        assert torch.allclose(rgb, torch.from_numpy(g[f"rgb{deg}"]), atol=2e-6, rtol=1e-6)
    assert np.allclose((np.array([0.0, 0.25, 1.0]) - 0.5) / O.SH_C0, g["RGB2SH"], atol=1e-6)
    assert np.allclose(np.array([-1.0, 0.0, 2.0]) * O.SH_C0 + 0.5, g["SH2RGB"], atol=1e-6)


def test_cov3d_matches_reference_build_scaling_rotation():
    g = np.load(os.path.join(GOLDEN, "ref_cov3d.npz"))
    cov = O.cov3d_from_scale_rot(torch.from_numpy(g["scales"]), torch.from_numpy(g["rotations"]), float(g["modifier"]))
    ref = torch.from_numpy(g["cov"])
    assert torch.allclose(cov, ref, rtol=2e-5, atol=1e-9)


def test_camera_matrices_match_reference_minicam():
    g = np.load(os.path.join(GOLDEN, "ref_camera.npz"))
    for i in range(3):
        W, H, fovy = g[f"dims{i}"]
        st = O.minicam_settings(g[f"c2w{i}"], int(W), int(H), float(fovy))
        assert np.allclose(st.viewmatrix.numpy(), g[f"wvt{i}"], atol=1e-6)
        assert np.allclose(st.projmatrix.numpy(), g[f"full{i}"], atol=1e-5)
        assert np.allclose(st.campos.numpy(), g[f"center{i}"], atol=1e-7)
        fy = np.deg2rad(fovy); fx = 2 * np.arctan(np.tan(fy / 2) * W / H)
        assert np.allclose(O.projection_matrix(0.01, 100.0, fx, fy).numpy(), g[f"proj{i}"], atol=1e-6)


def test_oracle_config0_regression_fixture():
    g = np.load(os.path.join(GOLDEN, "oracle_config0.npz"))
    st = _settings(128, 128)
    t = lambda k: torch.from_numpy(g[k])
    color, radii, depth, alpha, aux = O.rasterize(t("means3D"), None, t("shs"), None, t("opacities"), t("scales"),
                                                  t("rotations"), None, st, return_aux=True)
    assert np.array_equal(radii.numpy(), g["radii"])
    assert np.array_equal(aux["keys"], g["keys"]) and np.array_equal(aux["point_list"], g["point_list"])
    assert np.array_equal(aux["ranges"], g["ranges"])
    assert np.allclose(color.numpy(), g["color"], atol=1e-6) and np.allclose(alpha.numpy(), g["alpha"], atol=1e-6)


# ---------------- hand-computable micro cases (SURVEY §8c golden vectors (i)) ----------------------------
def _one(mean, scale=0.05, opacity=0.8, color=(1.0, 0.5, 0.25)):
    return dict(means3D=torch.tensor([mean], dtype=torch.float32), colors_precomp=torch.tensor([color]),
                opacities=torch.tensor([[opacity]]), scales=torch.full((1, 3), scale),
                rotations=torch.tensor([[1.0, 0, 0, 0]]))


def _render(c, st, aux=False):
    return O.rasterize(c["means3D"], None, None, c["colors_precomp"], c["opacities"], c["scales"], c["rotations"],
                       None, st, return_aux=aux)


def test_single_gaussian_on_axis():
    W = H = 64
    st = _settings(W, H)
    c = _one([0.0, 0.0, 0.0], scale=0.05, opacity=0.8)
    color, radii, depth, alpha, aux = _render(c, st, aux=True)
    pre = aux["pre"]
    z = 1.75
    f = W / (2 * st.tanfovx)
    expect = (f * 0.05 / z) ** 2 + 0.3                       # isotropic: cov2D = (f sigma / z)^2 + 0.3
    assert abs(float(pre["cov2d"][0, 0]) - expect) < 1e-3 * expect
    assert abs(float(pre["cov2d"][0, 2]) - expect) < 1e-3 * expect
    assert abs(float(pre["cov2d"][0, 1])) < 1e-4
    assert abs(float(pre["depth"][0]) - z) < 1e-5
    assert abs(float(pre["xy"][0, 0]) - (W - 1) / 2) < 1e-3   # pixel centres at integer coordinates
    assert int(radii[0]) == math.ceil(3 * math.sqrt(expect))
    # centre lies between the 4 middle pixels at distance sqrt(.5)
    a = 0.8 * math.exp(-0.5 * 0.5 / expect)
    assert abs(float(alpha[0, 31, 31]) - a) < 1e-5
    assert abs(float(color[0, 31, 31]) - a * 1.0) < 1e-5 and abs(float(depth[0, 31, 31]) - a * z) < 1e-4


def test_alpha_clamped_at_099():
    st = _settings(64, 64)
    c = _one([0.0, 0.0, 0.0], scale=0.3, opacity=1.0)
    _, _, _, alpha = _render(c, st)
    assert abs(float(alpha.max()) - 0.99) < 1e-6


def test_two_gaussians_order_and_transmittance():
    st = _settings(64, 64)
    near = _one([0.0, 0.0, 0.3], scale=0.2, opacity=0.6, color=(1, 0, 0))     # camera at +z (orbit az=0): larger z is nearer
    far = _one([0.0, 0.0, -0.3], scale=0.2, opacity=0.6, color=(0, 1, 0))
    c = {k: torch.cat([far[k], near[k]]) for k in near}
    color, _, _, alpha, aux = _render(c, st, aux=True)
    pre = aux["pre"]
    assert pre["depth"][1] < pre["depth"][0]
    y = x = 32
    tid = (y // 16) * 4 + x // 16
    s, e = aux["ranges"][tid]
    assert list(aux["point_list"][s:e]) == [1, 0]             # front to back
    def a_of(i):
        d = pre["xy"][i] - torch.tensor([float(x), float(y)])
        con = pre["conic"][i]
        p = -0.5 * (con[0] * d[0] ** 2 + con[2] * d[1] ** 2) - con[1] * d[0] * d[1]
        return float(min(0.99, 0.6 * math.exp(float(p))))
    a1, a0 = a_of(1), a_of(0)
    assert abs(float(color[0, y, x]) - a1) < 1e-5                        # red: nearest, T=1
    assert abs(float(color[1, y, x]) - a0 * (1 - a1)) < 1e-5             # green: behind, T=1-a1
    assert abs(float(alpha[0, y, x]) - (a1 + a0 * (1 - a1))) < 1e-5


def test_gaussian_straddling_tile_corner_emits_4_keys():
    st = _settings(64, 64)
    c = _one([0.0, 0.0, 0.0], scale=0.01, opacity=0.5)       # radius small, centre at pixel 31.5 -> tiles (1,1),(2,1),(1,2),(2,2)
    _, radii, _, _, aux = _render(c, st, aux=True)
    assert int(radii[0]) >= 1
    tiles = sorted(int(k >> np.uint64(32)) for k in aux["keys"])
    assert tiles == [1 * 4 + 1, 1 * 4 + 2, 2 * 4 + 1, 2 * 4 + 2]
    bits = np.float32(aux["pre"]["depth"][0].item()).view(np.uint32)
    assert all(np.uint32(k & np.uint64(0xFFFFFFFF)) == bits for k in aux["keys"])


def test_near_plane_cull_edge():
    st = _settings(64, 64)
    # camera at z=1.75 looking down -z: view depth = 1.75 - z_world; cull is depth <= 0.2 (fp32 compare)
    for zw in [1.75 - 0.15, 1.75 - 0.19999, 1.75 - 0.2, 1.75 - 0.20001, 1.75 - 0.25, 1.75 + 0.5]:
        c = _one([0.0, 0.0, zw], scale=0.001, opacity=0.5)
        _, radii, _, _, aux = _render(c, st, aux=True)
        depth = np.float32(aux["pre"]["depth"][0].item())
        assert (int(radii[0]) > 0) == bool(depth > np.float32(0.2)), (zw, depth, int(radii[0]))
    c = _one([0.0, 0.0, 1.75 - 0.15]); assert int(_render(c, st)[1][0]) == 0
    c = _one([0.0, 0.0, 1.75 - 0.25]); assert int(_render(c, st)[1][0]) > 0


def test_background_and_empty_scene():
    st = _settings(32, 32, bg=(0.2, 0.4, 0.6))
    z = torch.zeros
    color, radii, depth, alpha = O.rasterize(z(0, 3), None, None, z(0, 3), z(0, 1), z(0, 3), z(0, 4), None, st)
    assert radii.numel() == 0 and float(alpha.abs().max()) == 0 and float(depth.abs().max()) == 0
    assert torch.allclose(color[:, 5, 7], torch.tensor([0.2, 0.4, 0.6]))


def test_exclusive_argument_errors():
    st = _settings(32, 32)
    c = _one([0, 0, 0])
    with pytest.raises(Exception, match="excatly one of either SHs"):
        O.rasterize(c["means3D"], None, None, None, c["opacities"], c["scales"], c["rotations"], None, st)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed"):
        O.rasterize(c["means3D"], None, None, c["colors_precomp"], c["opacities"], None, None, None, st)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed"):
        O.rasterize(c["means3D"], None, None, c["colors_precomp"], c["opacities"], c["scales"], c["rotations"],
                    torch.zeros(1, 6), st)


# ---------------- vectorised composite == literal sequential recurrence -------------------------------------
@pytest.mark.parametrize("kind,deg,W,H", [("D0", 0, 64, 48), ("D1", 2, 80, 72)])
def test_composite_vectorised_equals_sequential(kind, deg, W, H):
    cl = O.make_cloud(kind, 600, deg, seed=3)
    st = _settings(W, H, el=15, az=40, deg=deg, bg=(0.1, 0.2, 0.3))
    color, radii, depth, alpha, aux = O.rasterize(cl["means3D"], None, cl["shs"], None, cl["opacities"], cl["scales"],
                                                  cl["rotations"], None, st, return_aux=True)
    c2, d2, a2, n2, T2 = O.composite_sequential(aux["pre"], aux["point_list"], aux["ranges"], st)
    assert float((c2 - color).abs().max()) < 5e-6 and float((d2 - depth).abs().max()) < 1e-5
    assert float((a2 - alpha).abs().max()) < 5e-6
    assert int((n2 != aux["n_contrib"]).sum()) == 0
    assert float((T2 - aux["final_T"]).abs().max()) < 1e-6
    # alpha_out + final_T == 1 (transmittance bookkeeping)
    assert float((alpha[0] + aux["final_T"] - 1).abs().max()) < 1e-5
    # keys sorted, ranges partition the list
    k = aux["keys"]
    assert np.all(k[:-1] <= k[1:])
    r = aux["ranges"].astype(np.int64)
    assert int((r[:, 1] - r[:, 0]).sum()) == k.size


@pytest.mark.parametrize("kind,n,W,H", [("D0", 3000, 200, 120), ("D1", 2500, 160, 96)])
def test_vectorised_key_duplication_equals_the_loop_and_tile_subset_composite(kind, n, W, H):
    """The two helpers the full-size GPU parity test relies on: keys/values without the per-Gaussian Python loop, and
    the composite restricted to a subset of tiles (identical pixels inside the subset, background outside)."""
    cl = O.make_cloud(kind, n, 1, seed=5)
    st = _settings(W, H, 5, 30, deg=1, bg=(0.1, 0.2, 0.3))
    pre = O.preprocess(cl["means3D"], cl["scales"], cl["rotations"], cl["opacities"], cl["shs"], None, None, None, st)
    k0, v0 = O.duplicate_with_keys(pre)
    k1, v1 = O.duplicate_with_keys_vectorised(pre)
    assert np.array_equal(k0, k1) and np.array_equal(v0, v1)
    sk, sv = O.sort_pairs(k1, v1)
    gx, gy = pre["grid"]
    ranges = O.identify_tile_ranges(sk, gx * gy)
    full = O.composite(pre, sv, ranges, st)
    subset = [t for t in range(gx * gy) if t % 3 == 1]
    part = O.composite(pre, sv, ranges, st, tiles=subset)
    mask = torch.zeros(H, W, dtype=torch.bool)
    for t in subset:
        yy, xx = O._tile_pixels(t, gx, W, H, torch.float32)
        mask[yy, xx] = True
    for a, b in zip(full[:3], part[:3]):
        assert torch.equal(a[:, mask], b[:, mask])
    assert torch.equal(full[3][mask], part[3][mask]) and torch.equal(full[4][mask], part[4][mask])
    assert torch.equal(part[0][:, ~mask], st.bg[:, None].expand(3, int((~mask).sum())).to(part[0].dtype))


def test_sort_is_stable_for_equal_keys():
    keys = np.array([5, 3, 5, 3, 5], dtype=np.uint64)
    vals = np.arange(5, dtype=np.uint32)
    sk, sv = O.sort_pairs(keys, vals)
    assert list(sv) == [1, 3, 0, 2, 4]


# ---------------- gradients: autograd vs central differences in fp64 ---------------------------------------
def test_gradients_match_finite_differences_fp64():
    torch.manual_seed(0)
    N, W, H, deg = 12, 32, 32, 2
    cl = O.make_cloud("D1", N, deg, seed=11)
    cl = {k: v.double() for k, v in cl.items()}
    cl["scales"] = cl["scales"] * 4.0
    st = _settings(W, H, el=10, az=25, deg=deg, bg=(0.3, 0.1, 0.2))
    g = torch.Generator().manual_seed(1)
    dc = (torch.rand(3, H, W, generator=g) * 2 - 1).double()
    dd = (torch.rand(1, H, W, generator=g) * 0.2 - 0.1).double()
    da = (torch.rand(1, H, W, generator=g) * 0.2 - 0.1).double()
    names = ("means3D", "shs", "opacities", "scales", "rotations")
    _, grads = O.rasterize_with_grads({k: cl[k] for k in names}, st, dc, dd, da)

    def loss(inp):
        c, _, d, a = O.rasterize(inp["means3D"], None, inp["shs"], None, inp["opacities"], inp["scales"],
                                 inp["rotations"], None, st)
        return float((c * dc).sum() + (d * dd).sum() + (a * da).sum())

    rng = np.random.RandomState(0)
    for name in names:
        flat = cl[name].reshape(-1)
        for j in rng.choice(flat.numel(), size=6, replace=False):
            eps = 1e-6
            p = {k: v.clone() for k, v in cl.items()}; m = {k: v.clone() for k, v in cl.items()}
            p[name].view(-1)[j] += eps; m[name].view(-1)[j] -= eps
            fd = (loss(p) - loss(m)) / (2 * eps)
            an = float(grads[name].reshape(-1)[j])
            assert abs(fd - an) <= 2e-4 * max(1.0, abs(fd)), (name, int(j), fd, an)


def test_means2d_grad_is_ndc_gradient_of_pixel_position():
    # means2D enters as an NDC offset: d(px)/d(ndc) = W/2  (the value densification reads, main_3DGS_renderer.py:767-769)
    cl = O.make_cloud("D1", 50, 0, seed=2)
    st = _settings(48, 32, az=10)
    H, W = 32, 48
    dc = torch.ones(3, H, W); dd = torch.zeros(1, H, W); da = torch.zeros(1, H, W)
    names = ("means3D", "shs", "opacities", "scales", "rotations")
    _, grads = O.rasterize_with_grads({k: cl[k] for k in names}, st, dc, dd, da)
    assert grads["means2D"].shape == (50, 3) and float(grads["means2D"][:, 2].abs().max()) == 0.0
    assert float(grads["means2D"][:, :2].abs().max()) > 0
