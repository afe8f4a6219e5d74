"""kiui.op.uv_padding shim (color_func_to_albedo, mesh_processer/mesh_utils.py:566): brute-force definition check."""
import numpy as np
import torch

from conftest import ROOT  # noqa: F401
from kiui.op import uv_padding


def _brute(img, mask, p):
    H, W = mask.shape
    ys, xs = np.nonzero(mask)
    out = img.copy()
    amb = np.zeros((H, W), bool)
    for y in range(H):
        for x in range(W):
            if mask[y, x]:
                continue
            if (np.abs(ys - y) + np.abs(xs - x)).min() > p:      # outside p steps of 4-connected dilation
                continue
            d2 = (ys - y) ** 2 + (xs - x) ** 2
            k = np.flatnonzero(d2 == d2.min())
            amb[y, x] = len(k) > 1                               # ties: the kd-tree's pick is implementation defined
            out[y, x] = img[ys[k[0]], xs[k[0]]]
    return out, amb


def test_uv_padding_matches_definition_small_and_kdtree_paths():
    g = torch.Generator().manual_seed(0)
    H, W = 40, 52
    img = torch.rand(H, W, 3, generator=g)
    mask = torch.zeros(H, W, dtype=torch.bool)
    mask[8:20, 10:30] = True; mask[25:33, 35:47] = True; mask[30, 5] = True
    for p in (1, 2, 5, 11):                                     # 11 > window limit -> kd-tree path
        ref, amb = _brute(img.numpy(), mask.numpy(), p)
        got = uv_padding(img, mask, p)
        assert torch.is_tensor(got) and got.shape == img.shape
        assert torch.equal(got[mask], img[mask])                 # valid texels untouched
        ok = ~torch.from_numpy(amb)
        assert np.allclose(got.numpy()[ok.numpy()], ref[ok.numpy()]), p
        # ambiguous texels still hold the colour of SOME nearest valid texel: they changed from the original
        far = torch.from_numpy((ref == img.numpy()).all(-1)) & ~mask
        assert torch.equal(got[far], img[far])                   # beyond the padding: untouched
    # ndarray in -> ndarray out; default padding = 0.1 * max(H, W)
    out = uv_padding(img.numpy(), mask.numpy())
    assert isinstance(out, np.ndarray) and out.shape == (H, W, 3)
    # degenerate masks
    assert torch.equal(uv_padding(img, torch.zeros(H, W, dtype=torch.bool), 2), img)
    assert torch.equal(uv_padding(img, torch.ones(H, W, dtype=torch.bool), 2), img)


def test_mesh_regularisers_on_known_meshes():
    from kiui.mesh_utils import laplacian_smooth_loss, normal_consistency, clean_mesh
    from kiui.op import dot
    import pytest
    # flat 3x3 grid of vertices (8 triangles): interior vertex has zero uniform-Laplacian residual, all normals agree
    ys, xs = torch.meshgrid(torch.arange(3.0), torch.arange(3.0), indexing="ij")
    v = torch.stack([xs.flatten(), ys.flatten(), torch.zeros(9)], 1).requires_grad_(True)
    f = []
    for y in range(2):
        for x in range(2):
            a = y * 3 + x
            f += [[a, a + 1, a + 4], [a, a + 4, a + 3]]
    f = torch.tensor(f, dtype=torch.int32)
    assert float(normal_consistency(v, f)) < 1e-7
    lap = laplacian_smooth_loss(v, f)
    # manual: L v = deg_i * v_i - sum_{j~i} v_j
    nb = {i: set() for i in range(9)}
    for t in f.tolist():
        for a, b in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0])):
            nb[a].add(b); nb[b].add(a)
    man = torch.stack([len(nb[i]) * v[i] - sum(v[j] for j in nb[i]) for i in range(9)]).norm(dim=1).mean()
    assert abs(float(lap) - float(man)) < 1e-6
    lap.backward(); assert torch.isfinite(v.grad).all()
    # regular tetrahedron: every edge has dihedral cos(normals) = -1/3 -> |1 - cos| = 4/3
    tv = torch.tensor([[1.0, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]])
    tf = torch.tensor([[0, 1, 2], [0, 3, 1], [0, 2, 3], [1, 3, 2]])
    assert abs(float(normal_consistency(tv, tf)) - 4.0 / 3.0) < 1e-6
    assert dot(tv, tv).shape == (4, 1) and float(dot(tv, tv)[0]) == 3.0
    with pytest.raises(NotImplementedError):
        clean_mesh(tv, tf)
