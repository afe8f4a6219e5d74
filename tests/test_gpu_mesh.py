"""GPU parity tests of the mesh path (nvdiffrast surface) against oracle/dr_oracle.py, through the C ABI
(include/dr_b200.h) via the `nvdiffrast.torch` shim.  Triangle ids bit-exact, floats within tolerance, gradients
within 1e-3 relative."""
import numpy as np
import pytest
import torch

from conftest import ROOT  # noqa: F401
from oracle import dr_oracle as D
from oracle import gs_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def dr():
    import nvdiffrast.torch as dr
    return dr


def _scene(subdiv, H, W, views, seed=0, jitter=0.0):
    v, f, uv = D.icosphere(subdiv)
    if jitter:
        v = v + jitter * torch.randn(v.shape, generator=torch.Generator().manual_seed(seed))
    proj = D.gl_perspective(49.1, W / H)
    pos = torch.cat([D.clip_positions(v, O.orbit_camera(el, az, 1.75), proj) for el, az in views], dim=0)
    return v, f, uv, pos


def _rel(a, b):
    return float((a.detach().cpu().double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.mark.parametrize("subdiv,H,W,views,jit", [(2, 64, 64, [(20, 30)], 0.0), (3, 96, 128, [(0, 0), (-30, 140)], 0.02),
                                                  (4, 120, 200, [(10, 75), (40, -100), (-15, 10)], 0.0)])
def test_rasterize_ids_bit_exact_and_values(dr, dev, subdiv, H, W, views, jit):
    v, f, uv, pos = _scene(subdiv, H, W, views, jitter=jit)
    ref, ref_db = D.rasterize(pos, f, (H, W))
    rast, db = dr.rasterize(dr.RasterizeCudaContext(), pos.to(dev), f.to(dev), (H, W))
    assert rast.shape == (len(views), H, W, 4) and db.shape == rast.shape
    assert torch.equal(rast[..., 3].cpu(), ref[..., 3])                       # triangle ids: bit-exact
    assert float((rast[..., :3].cpu() - ref[..., :3]).abs().max()) < 2e-5
    assert float((db.cpu() - ref_db).abs().max()) < 2e-4 * max(1.0, float(ref_db.abs().max()))
    r2, _ = dr.rasterize(dr.RasterizeGLContext(), pos.to(dev), f.to(dev), (H, W))
    assert torch.equal(r2, rast)                                               # deterministic


def test_full_render_chain_forward_and_gradients(dr, dev):
    """The op order of DiffRastRenderer.render (diff_mesh_renderer.py:94-139) with train_geo on."""
    H, W = 96, 96
    views = [(15, 40), (-25, 200)]
    v, f, uv, _ = _scene(3, H, W, views)
    proj = D.gl_perspective(49.1, W / H)
    tex0 = torch.rand(1, 32, 32, 3, generator=torch.Generator().manual_seed(1))
    gi = torch.rand(2, H, W, 3, generator=torch.Generator().manual_seed(2)); ga = torch.rand(2, H, W, 1, generator=torch.Generator().manual_seed(3))
    gd = torch.rand(2, H, W, 1, generator=torch.Generator().manual_seed(4))

    def chain(ops, rasterize, vv, tex, device):
        pos = torch.cat([D.clip_positions(vv.cpu(), O.orbit_camera(el, az, 1.75), proj) if False else
                         (torch.nn.functional.pad(vv, (0, 1), value=1.0) @ torch.inverse(torch.from_numpy(O.orbit_camera(el, az, 1.75))).T.to(device)
                          @ torch.from_numpy(proj).T.to(device))[None] for el, az in views], dim=0)
        ff = f.to(device)
        rast, db = rasterize(pos, ff)
        alpha = ops.antialias(torch.clamp(rast[..., -1:], 0, 1).contiguous(), rast, pos, ff).clamp(0, 1)
        texc, texc_db = ops.interpolate(uv.to(device)[None].contiguous(), rast, ff, rast_db=db, diff_attrs="all")
        albedo = torch.sigmoid(ops.texture(tex, texc)) if ops is D else torch.sigmoid(ops.texture(tex, texc, uv_da=texc_db, filter_mode="linear"))
        depth, _ = ops.interpolate(pos[..., 2:3].contiguous(), rast, ff)
        albedo = ops.antialias(albedo, rast, pos, ff)
        img = alpha * albedo + (1 - alpha) * 1.0
        return img, alpha, depth

    vr = v.clone().requires_grad_(True); tr = tex0.clone().requires_grad_(True)
    img_r, a_r, d_r = chain(D, lambda p, ff: D.rasterize(p, ff, (H, W)), vr, tr, "cpu")
    ((img_r * gi).sum() + (a_r * ga).sum() + (d_r * gd).sum()).backward()
    vg = v.to(dev).requires_grad_(True); tg = tex0.to(dev).requires_grad_(True)
    ctx = dr.RasterizeCudaContext()
    img_g, a_g, d_g = chain(dr, lambda p, ff: dr.rasterize(ctx, p, ff, (H, W)), vg, tg, dev)
    ((img_g * gi.to(dev)).sum() + (a_g * ga.to(dev)).sum() + (d_g * gd.to(dev)).sum()).backward()
    assert float((img_g.detach().cpu() - img_r.detach()).abs().max()) < 1e-4
    assert float((a_g.detach().cpu() - a_r.detach()).abs().max()) < 1e-4
    assert float((d_g.detach().cpu() - d_r.detach()).abs().max()) < 1e-4
    assert _rel(tg.grad, tr.grad) < 1e-3
    assert _rel(vg.grad, vr.grad) < 1e-3


def test_interpolate_attr_batch_and_no_derivatives(dr, dev):
    H, W = 48, 64
    v, f, uv, pos = _scene(2, H, W, [(0, 0), (30, 90)])
    rast_r, db_r = D.rasterize(pos, f, (H, W))
    rast, db = dr.rasterize(dr.RasterizeCudaContext(), pos.to(dev), f.to(dev), (H, W))
    attr_b = torch.rand(2, v.shape[0], 5, generator=torch.Generator().manual_seed(0))
    for attr in (attr_b, attr_b[:1]):
        ar = attr.clone().requires_grad_(True); ag = attr.to(dev).requires_grad_(True)
        o_r, da_r = D.interpolate(ar, rast_r, f, rast_db=db_r, diff_attrs="all")
        o_g, da_g = dr.interpolate(ag, rast, f.to(dev), rast_db=db, diff_attrs="all")
        assert float((o_g.detach().cpu() - o_r.detach()).abs().max()) < 1e-5
        assert float((da_g.detach().cpu() - da_r.detach()).abs().max()) < 1e-4
        w1 = torch.rand_like(o_r); w2 = torch.rand_like(da_r)
        ((o_r * w1).sum() + (da_r * w2).sum()).backward()
        ((o_g * w1.to(dev)).sum() + (da_g * w2.to(dev)).sum()).backward()
        assert _rel(ag.grad, ar.grad) < 1e-4
    o_g, none = dr.interpolate(attr_b.to(dev), rast, f.to(dev))
    assert none is None and o_g.shape == (2, H, W, 5)


def test_texture_wrap_clamp_and_gradients(dr, dev):
    g = torch.Generator().manual_seed(0)
    tex = torch.rand(1, 16, 24, 4, generator=g)
    uv = torch.rand(2, 20, 30, 2, generator=g) * 3 - 1          # outside [0,1] exercises wrap / clamp
    for mode in ("wrap", "clamp"):
        tr = tex.clone().requires_grad_(True); ur = uv.clone().requires_grad_(True)
        tg = tex.to(dev).requires_grad_(True); ug = uv.to(dev).requires_grad_(True)
        o_r = D.texture(tr, ur, boundary_mode=mode)
        o_g = dr.texture(tg, ug, filter_mode="linear", boundary_mode=mode)
        assert float((o_g.detach().cpu() - o_r.detach()).abs().max()) < 1e-5
        wgt = torch.rand_like(o_r)
        (o_r * wgt).sum().backward(); (o_g * wgt.to(dev)).sum().backward()
        assert _rel(tg.grad, tr.grad) < 1e-4 and _rel(ug.grad, ur.grad) < 1e-3


def test_topology_matches_oracle(dev):
    from gs_b200 import meshops
    for sub in (0, 2, 3):
        v, f, _ = D.icosphere(sub)
        opp = meshops.get_topology(f.to(dev), v.shape[0])
        assert torch.equal(opp.cpu().to(torch.int64), D.edge_opposites(f))
    f_open = D.icosphere(2)[1][:50].contiguous()                  # open mesh: boundary edges -> -1
    opp = meshops.get_topology(f_open.to(dev), 162)
    assert torch.equal(opp.cpu().to(torch.int64), D.edge_opposites(f_open))


def test_uv_space_bake_like_color_func_to_albedo(dr, dev):
    """mesh_utils.py:521-541: rasterize in UV space (uv*2-1, z=0, w=1) and interpolate xyz."""
    v, f, uv = D.icosphere(2)
    pos = torch.cat([uv * 2 - 1, torch.zeros_like(uv[:, :1]), torch.ones_like(uv[:, :1])], dim=-1)[None]
    ref, _ = D.rasterize(pos, f, (128, 128))
    rast, _ = dr.rasterize(dr.RasterizeCudaContext(), pos.to(dev), f.to(dev), (128, 128))
    assert torch.equal(rast[..., 3].cpu(), ref[..., 3])
    xyz_r, _ = D.interpolate(v[None], ref, f)
    xyz_g, _ = dr.interpolate(v.to(dev)[None], rast, f.to(dev))
    assert float((xyz_g.cpu() - xyz_r).abs().max()) < 1e-5


def test_full_size_properties_config4(dr, dev):
    """BASELINE config 4 scale: 327,680-triangle sphere, 8 views at 1920x1080 (the 500k-triangle class)."""
    H, W = 1080, 1920
    views = [(0, 45.0 * k) for k in range(8)]
    v, f, uv, pos = _scene(7, H, W, views)
    assert f.shape[0] == 327680
    posd, fd = pos.to(dev), f.to(dev)
    ctx = dr.RasterizeCudaContext()
    rast, db = dr.rasterize(ctx, posd, fd, (H, W))
    ids = rast[..., 3]
    hit = ids > 0
    assert bool((ids[hit] <= f.shape[0]).all()) and bool((rast[..., 2][hit].abs() <= 1).all())
    u, vv = rast[..., 0][hit], rast[..., 1][hit]
    assert float(u.min()) >= -1e-5 and float(vv.min()) >= -1e-5 and float((u + vv).max()) <= 1 + 1e-5
    # watertight: the projected sphere's interior has no holes (disc of radius 0.95*r around the centre)
    cov = hit[0].float()
    ys, xs = torch.nonzero(hit[0], as_tuple=True)
    cy, cx = ys.float().mean(), xs.float().mean()
    r = 0.5 * (xs.max() - xs.min()).float() * 0.95
    yy, xx = torch.meshgrid(torch.arange(H, device=dev).float(), torch.arange(W, device=dev).float(), indexing="ij")
    disc = ((yy - cy) ** 2 + (xx - cx) ** 2) < r * r
    assert bool(hit[0][disc].all())
    r2, _ = dr.rasterize(ctx, posd, fd, (H, W))
    assert torch.equal(r2, rast)                                   # order-independent 64-bit atomicMin -> deterministic
    # antialias only touches pixels adjacent to an id discontinuity and keeps values in range
    alpha = torch.clamp(ids[..., None], 0, 1).contiguous()
    aa = dr.antialias(alpha, rast, posd, fd)
    changed = (aa - alpha).abs() > 0
    nb = torch.zeros_like(hit)
    nb[:, :, 1:] |= ids[:, :, 1:] != ids[:, :, :-1]; nb[:, :, :-1] |= ids[:, :, 1:] != ids[:, :, :-1]
    nb[:, 1:, :] |= ids[:, 1:, :] != ids[:, :-1, :]; nb[:, :-1, :] |= ids[:, 1:, :] != ids[:, :-1, :]
    assert bool((~changed[..., 0] | nb).all()) and float(aa.min()) >= -1e-6 and float(aa.max()) <= 1 + 1e-6
    assert int(changed.sum()) > 0          # (with ~1.5 px triangles few boundary pixels own a crossing silhouette edge)


# ---------------- FlexiCubesRenderer.render_mesh on the CUDA shim (rows c2 / c4) ------------------------------------------
def _render_mesh_like_the_reference(dr, v, f, mv, mvp, res, dev):
    """FlexiCubesRenderer.render_mesh (MVs_Algorithms/FlexiCubes/flexicubes_renderer.py:40-74) written against the shim:
    a FRESH RasterizeCudaContext per call (:46), batched views, util.interpolate with per-face attribute indices for the
    face normals (:65-66), antialias of the mask and of the vertex normals, white background."""
    v, f = v.to(dev), f.to(dev)
    hom = torch.nn.functional.pad(v[None], (0, 1), value=1.0)
    pos = torch.matmul(hom, mvp.to(dev).transpose(1, 2))                        # util.xfm_points
    fi = f.int()
    rast, _ = dr.rasterize(dr.RasterizeCudaContext(), pos, fi, res)
    a_idx = rast[..., -1:] > 0
    alpha = a_idx.float()
    white = lambda img: torch.lerp(torch.ones_like(img), img, alpha)
    out = {"mask": white(dr.antialias(alpha, rast, pos, fi))}
    v_cam = torch.matmul(hom, mv.to(dev).transpose(1, 2))
    depth, _ = dr.interpolate(v_cam[..., [2]].contiguous(), rast, fi)
    dm = torch.clamp(depth[a_idx], min=-5.5, max=-0.5)
    depth[a_idx] = (dm + 5.5) / 5.0
    depth[~a_idx] = 0
    out["depth"] = white(depth)
    fl = f.long()
    fn = torch.cross(v[fl[:, 1]] - v[fl[:, 0]], v[fl[:, 2]] - v[fl[:, 0]], dim=-1)
    fn = fn / torch.sqrt(torch.clamp((fn * fn).sum(-1, keepdim=True), min=1e-20))          # util.safe_normalize
    nidx = torch.arange(fn.shape[0], dtype=torch.int64, device=dev)[:, None].repeat(1, 3)
    normal, _ = dr.interpolate(fn[None].contiguous(), rast, nidx.int())
    out["normal"] = white(normal)
    vn = v / torch.sqrt(torch.clamp((v * v).sum(-1, keepdim=True), min=1e-20))
    vnorm, _ = dr.interpolate(vn[None].contiguous(), rast, fi)
    out["vertex_normal"] = white(dr.antialias((vnorm + 1) * 0.5, rast, pos, fi))
    return out


@pytest.mark.parametrize("which", ["icosphere", "flexicubes_extraction"])
def test_flexicubes_render_mesh_on_the_cuda_shim_matches_the_reference_renderer(dr, dev, which):
    """tests/golden/ref_host.npz holds what the REFERENCE classes produced on the CPU (make_golden_training.py:
    FlexiCubesRenderer.render_mesh from the reference source over the mesh oracle) for (a) an icosphere and (b) the mesh
    the reference's own FlexiCubes.__call__ (flexicubes.py:133-216) extracted from a bumpy-sphere SDF on a 14^3 grid.
    The same call sequence on the CUDA shim must reproduce those images: one silhouette pixel may flip a tie, everything
    else within 1e-4."""
    import os
    from conftest import GOLDEN
    Hh = np.load(os.path.join(GOLDEN, "ref_host.npz"))
    mv, mvp = torch.from_numpy(Hh["flexi_mv"]), torch.from_numpy(Hh["flexi_mvp"])
    res = (int(Hh["flexi_res"][0]), int(Hh["flexi_res"][1]))
    if which == "icosphere":
        v, f, _ = D.icosphere(2, 0.8)
        pre = "flexi_"
    else:
        v, f = torch.from_numpy(Hh["flexi_ex_v"]), torch.from_numpy(Hh["flexi_ex_f"])
        pre = "flexi_ex_"
        assert f.shape[0] > 500 and int(f.max()) < v.shape[0]
    for _ in range(2):               # a second call with a second fresh context: scratch is cached per device, results identical
        out = _render_mesh_like_the_reference(dr, v, f, mv, mvp, res, dev)
    for k in ("mask", "depth", "normal", "vertex_normal"):
        got, ref = out[k].cpu().numpy(), Hh[pre + k]
        assert got.shape == ref.shape, k
        err = np.abs(got - ref).max(axis=-1)
        assert int((err > 1e-4).sum()) <= 2, (k, int((err > 1e-4).sum()), float(err.max()))
    assert float(np.abs(Hh[pre + "normal"][0] - Hh[pre + "normal"][1]).mean()) > 0.01       # the two views differ


def test_mipmapped_texture_auto_mode_like_lgm_render_mesh(dr, dev):
    """LGM's texture fit (Gen_3D_Modules/LGM/nerf_marching_cubes_converter.py:229-230): interpolate(uv, rast_db,
    diff_attrs='all') -> dr.texture(albedo, texc, uv_da=texc_db) with the DEFAULT filter_mode ('auto' -> trilinear mip
    mapping).  A far camera makes the footprint several texels wide, so levels > 0 are in play.  Values and gradients
    (texture, uv through the vertex uv attribute) against the oracle."""
    H = W = 96
    v, f, uv, _ = _scene(3, H, W, [(10, 30)])
    proj = D.gl_perspective(49.1, 1.0)
    pos = D.clip_positions(v, O.orbit_camera(10, 30, 4.0), proj)                # small on screen: minification
    tex0 = torch.rand(1, 128, 128, 3, generator=torch.Generator().manual_seed(3))
    gi = torch.rand(1, H, W, 3, generator=torch.Generator().manual_seed(4))

    def run(mod, dev_):
        tex = tex0.clone().to(dev_).requires_grad_(True)
        uva = uv.clone().to(dev_).requires_grad_(True)
        p, ff = pos.to(dev_), f.to(dev_)
        if mod is dr:
            rast, db = mod.rasterize(mod.RasterizeCudaContext(), p, ff, (H, W))
        else:
            rast, db = mod.rasterize(p, ff, (H, W))
        texc, texc_db = mod.interpolate(uva[None], rast, ff, rast_db=db, diff_attrs="all")
        if mod is dr:
            img = mod.texture(tex, texc, uv_da=texc_db)                                  # filter_mode defaults to 'auto'
        else:
            img = mod.texture(tex, texc, "linear-mipmap-linear", uv_da=texc_db)
        (img * gi.to(dev_)).sum().backward()
        return img.detach().cpu(), tex.grad.cpu(), uva.grad.cpu(), texc_db.detach().cpu()
    img, gt, gu, da = run(dr, dev)
    rimg, rgt, rgu, rda = run(D, torch.device("cpu"))
    lod = D.mip_level(rda, 128, 128, 8)
    assert float(lod.max()) > 1.5 and float((lod > 0.5).float().mean()) > 0.01          # mip levels really are used
    cover = (rimg.abs().sum(-1) > 0)
    assert float((img - rimg).abs()[cover].max()) < 2e-4
    assert _rel(gt, rgt) < 1e-3 and _rel(gu, rgu) < 2e-3
    # plain bilinear differs visibly here: the mode matters
    lin = dr.texture(tex0.to(dev), D.interpolate(uv[None], D.rasterize(pos, f, (H, W))[0], f)[0].to(dev), filter_mode="linear").cpu()
    assert float((lin - rimg).abs()[cover].mean()) > 5 * float((img - rimg).abs()[cover].mean()) + 1e-4
