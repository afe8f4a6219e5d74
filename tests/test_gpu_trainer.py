"""GPU tests of the optimisation-step kernels and the native trainer (rows a6/a8/a9):
fused activation, fused chain-rule+Adam vs torch.optim.Adam, densification statistics and rules, a short fit."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import ROOT  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def test_activate_matches_reference_activations(dev):
    from gs_b200 import _lib
    N = 5000
    g = torch.Generator(device="cpu").manual_seed(0)
    ro, rs, rr = torch.randn(N, 1, generator=g).to(dev), torch.randn(N, 3, generator=g).to(dev), torch.randn(N, 4, generator=g).to(dev)
    o, s, r = torch.empty_like(ro), torch.empty_like(rs), torch.empty_like(rr)
    _lib.check(_lib.lib.gs_b200_activate(N, _ptr(ro), _ptr(rs), _ptr(rr), _ptr(o), _ptr(s), _ptr(r), None))
    assert torch.allclose(o, torch.sigmoid(ro), atol=1e-6) and torch.allclose(s, torch.exp(rs), rtol=1e-6)
    assert torch.allclose(r, torch.nn.functional.normalize(rr), atol=1e-6)


def test_fused_adam_matches_torch_adam_through_activations(dev):
    """training_setup (main_3DGS_renderer.py:435-453): six groups, eps=1e-15; gradients arrive wrt ACTIVATED values."""
    from gs_b200 import _lib
    N, M = 3000, 4
    g = torch.Generator(device="cpu").manual_seed(1)
    sizes = [3 * N, 3 * M * N, N, 3 * N, 4 * N]
    raw = torch.randn(sum(sizes), generator=g).to(dev) * 0.5
    m1 = torch.zeros_like(raw); m2 = torch.zeros_like(raw)
    lrs = np.array([1.6e-3, 2.5e-3, 1.25e-4, 0.05, 0.005, 0.001], dtype=np.float32)

    def split(buf):
        o = 0; out = []
        for s, shp in zip(sizes, [(N, 3), (N, M, 3), (N, 1), (N, 3), (N, 4)]):
            out.append(buf[o:o + s].view(*shp)); o += s
        return out
    xyz, shs, op, sc, ro = [t.clone().requires_grad_(True) for t in split(raw)]
    dc = shs.detach()[:, :1].clone().requires_grad_(True); rest = shs.detach()[:, 1:].clone().requires_grad_(True)
    opt = torch.optim.Adam([{"params": [xyz], "lr": float(lrs[0])}, {"params": [dc], "lr": float(lrs[1])}, {"params": [rest], "lr": float(lrs[2])},
                            {"params": [op], "lr": float(lrs[3])}, {"params": [sc], "lr": float(lrs[4])}, {"params": [ro], "lr": float(lrs[5])}],
                           lr=0.0, eps=1e-15)
    for step in range(1, 4):
        gact = torch.randn(sum(sizes), generator=g).to(dev)            # dL/d(activated) packed
        gx, gs_, go, gsc, gr = split(gact)
        opt.zero_grad()
        act = (xyz * gx).sum() + (torch.cat([dc, rest], 1) * gs_).sum() + (torch.sigmoid(op) * go).sum() + \
              (torch.exp(sc) * gsc).sum() + (torch.nn.functional.normalize(ro) * gr).sum()
        act.backward(); opt.step()
        _lib.check(_lib.lib.gs_b200_adam_step(N, M, C.c_void_p(lrs.ctypes.data), 0.9, 0.999, 1e-15, step, 1.0, _ptr(gact), _ptr(raw),
                                              _ptr(m1), _ptr(m2), None))
        ref = torch.cat([xyz.detach().reshape(-1), torch.cat([dc, rest], 1).detach().reshape(-1), op.detach().reshape(-1),
                         sc.detach().reshape(-1), ro.detach().reshape(-1)])
        assert float((raw - ref).abs().max()) < 2e-5, step


def test_densify_stats(dev):
    from gs_b200 import _lib
    N = 1000
    g2 = torch.randn(N, 3, device=dev); radii = torch.randint(-1, 5, (N,), dtype=torch.int32, device=dev)
    acc = torch.rand(N, device=dev); den = torch.rand(N, device=dev).round(); mr = torch.rand(N, device=dev) * 3
    a0, d0, m0 = acc.clone(), den.clone(), mr.clone()
    _lib.check(_lib.lib.gs_b200_densify_stats(N, _ptr(g2), _ptr(radii), _ptr(acc), _ptr(den), _ptr(mr), None))
    vis = radii > 0
    assert torch.allclose(acc, torch.where(vis, a0 + g2[:, :2].norm(dim=1), a0), atol=1e-6)
    assert torch.equal(den, torch.where(vis, d0 + 1, d0)) and torch.equal(mr, torch.where(vis, torch.maximum(m0, radii.float()), m0))


def test_densify_and_prune_rules(dev):
    """Clone / split / prune decisions follow main_3DGS_renderer.py:641-668,752-781 on a seeded state."""
    from gs_b200 import trainer
    tr = trainer.GaussianTrainer(trainer.TrainParams(num_pts=2000, sh_degree=1), device=dev, seed=3)
    N = tr.N
    g = torch.Generator(device="cpu").manual_seed(0)
    tr.grad_accum.copy_(torch.rand(N, generator=g).to(dev) * 4e-4); tr.denom.fill_(1.0); tr.denom[:50] = 0.0       # NaN grads -> 0
    tr.v["scaling"].copy_(torch.log(torch.rand(N, 3, generator=g).to(dev) * 0.08 + 1e-3))                          # some > 0.01*4, some > 0.4? no
    tr.v["scaling"][:20] = float(np.log(0.5))                                                                        # too big in world space -> pruned
    tr.v["opacity"][100:140] = -8.0                                                                                  # opacity < 0.005 -> pruned
    tr.max_radii2D[200:230] = 5.0                                                                                    # big on screen -> pruned
    grads = tr.grad_accum / tr.denom; grads[grads.isnan()] = 0
    scal = torch.exp(tr.v["scaling"]).max(1).values
    big = scal > 0.01 * 4.0
    n_clone = int(((grads >= 2e-4) & ~big).sum()); n_split = int(((grads >= 2e-4) & big).sum())
    opac = torch.sigmoid(tr.v["opacity"]).squeeze(1)
    pruned_old = ((grads >= 2e-4) & big) | (opac < 0.005) | (tr.max_radii2D > 1.0) | (scal > 0.4)
    xyz_before = tr.v["xyz"].clone(); m1_before = tr._views(tr.m1, N)["xyz"].clone()
    info = tr.densify_and_prune(2e-4, 0.005, 4.0, 1.0, generator=None)
    assert info["cloned"] == n_clone and info["split"] == n_split and info["pruned"] == int(pruned_old.sum())
    kept = int((~pruned_old).sum())
    assert tr.N <= kept + n_clone + 2 * n_split and tr.N >= kept
    assert torch.equal(tr.v["xyz"][:kept], xyz_before[~pruned_old])                  # survivors keep order and values
    assert float(tr._views(tr.m1, tr.N)["xyz"][kept:].abs().max()) == 0.0            # new points start with zero Adam state
    assert float(tr.grad_accum.abs().max()) == 0 and tr.radii.numel() == tr.N
    tr.reset_opacity()
    assert float(torch.sigmoid(tr.v["opacity"]).max()) <= 0.01 + 1e-6


def test_short_fit_reduces_loss(dev):
    """End-to-end: render a target from a 'ground-truth' cloud, fit a fresh random cloud for a few steps."""
    from gs_b200 import camera, trainer
    W = H = 192
    V = 4
    views = camera.orbit_views(V, W, H)
    gt = trainer.GaussianTrainer(trainer.TrainParams(num_pts=3000, sh_degree=1), device=dev, seed=7)
    gt.v["shs"][:, 0, :] = torch.rand(gt.N, 3, device=dev) * 2 - 0.5
    gt.v["opacity"].fill_(1.5)
    gt.v["scaling"].add_(0.7)
    target = gt.render_views(views, W, H)
    ref_img = target[:, :3].clamp(0, 1).contiguous(); ref_mask = (target[:, 4:5] > 0.5).float()
    tr = trainer.GaussianTrainer(trainer.TrainParams(num_pts=3000, sh_degree=1, density_start_iter=10 ** 9), device=dev, seed=11)
    ls = [tr.train_step(views, W, H, ref_img, ref_mask) for _ in range(40)]
    assert all(np.isfinite(ls)) and ls[-1] < 0.8 * ls[0], (ls[0], ls[-1])
    assert tr.step_count == 40 and bool(torch.isfinite(tr.raw).all())
