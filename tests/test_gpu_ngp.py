"""GPU parity tests of the Instant-NGP path against oracle/ngp_oracle.py, through the C ABI (include/ngp_b200.h)
via the kiui / nerfacc shims.  Packed sample lists bit-exact; features, weights, gradients within tolerance."""
import numpy as np
import pytest
import torch

from conftest import ROOT  # noqa: F401
from oracle import gs_oracle as O
from oracle import ngp_oracle as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _rel(a, b):
    return float((a.detach().cpu().double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.mark.parametrize("L", [4, 12, 16])
def test_grid_encode_forward_backward_and_tv(dev, L):
    from kiui.gridencoder import GridEncoder
    torch.manual_seed(L)
    enc = GridEncoder(num_levels=L).to(dev)
    enc.embeddings.data.uniform_(-1, 1)
    off = G.grid_offsets(num_levels=L)
    assert np.array_equal(off, enc._offsets_np.astype(np.int64))
    x = torch.rand(5000, 3) * 2 - 1
    x[:4] = torch.tensor([[-1.0, -1, -1], [1, 1, 1], [0, 0, 0], [1, -1, 0.5]])        # box corners / faces
    emb = enc.embeddings.detach().cpu()
    ref = G.grid_encode(x, emb, off, num_levels=L)
    out = enc(x.to(dev))
    assert out.shape == (5000, 2 * L)
    assert float((out.detach().cpu() - ref).abs().max()) < 2e-5
    gout = torch.rand(5000, 2 * L)
    e = emb.clone().requires_grad_(True)
    (G.grid_encode(x, e, off, num_levels=L) * gout).sum().backward()
    (out * gout.to(dev)).sum().backward()
    assert _rel(enc.embeddings.grad, e.grad) < 1e-5
    # prefix shapes and bound
    o2 = enc(x.to(dev).view(50, 100, 3) * 0.5, bound=0.5)
    assert o2.shape == (50, 100, 2 * L) and float((o2.reshape(-1, 2 * L) - out).abs().max()) < 1e-6
    # total-variation gradient, explicit sample points
    xs = torch.rand(3000, 3) * 2 - 1
    g0 = enc.embeddings.grad.clone()
    enc.grad_total_variation(1e-3, inputs=xs.to(dev))
    ref_tv = G.grad_total_variation(xs, emb, off, 1e-3, num_levels=L)
    assert float(((enc.embeddings.grad - g0).cpu() - ref_tv).abs().max()) < 1e-6
    enc.grad_total_variation(1e-8)                              # the reference's call form (Instant_NGP.py:195)


def _binary_ball(R, radius):
    g = (torch.arange(R).float() + 0.5) / R * 2 - 1
    x, y, z = torch.meshgrid(g, g, g, indexing="ij")
    return (x * x + y * y + z * z) < radius * radius


@pytest.mark.parametrize("R,h,w,strat", [(64, 24, 32, False), (32, 40, 40, True)])
def test_marching_sample_list_bit_exact(dev, R, h, w, strat):
    from gs_b200 import ngp
    binary = _binary_ball(R, 0.6)
    aabb = torch.tensor([-1.0, -1, -1, 1, 1, 1])
    ro, rd = G.get_rays(O.orbit_camera(15, 40, 1.75), h, w, 49.1)
    toff = torch.rand(h * w, generator=torch.Generator().manual_seed(1)) * 5e-3 if strat else None
    ri_r, ts_r, te_r = G.march(ro, rd, binary, aabb, 0.01, 100.0, 5e-3, toff)
    ri, ts, te = ngp.march_rays(ro.to(dev), rd.to(dev), binary.to(dev), aabb.to(dev), 0.01, 100.0, 5e-3,
                                None if toff is None else toff.to(dev))
    assert ri.dtype == torch.int64 and ri.numel() == ri_r.numel() > 1000
    assert torch.equal(ri.cpu(), ri_r)
    assert torch.equal(ts.cpu(), ts_r) and torch.equal(te.cpu(), te_r)           # bit-exact floats by construction
    assert bool((ri[1:] >= ri[:-1]).all())


def test_weights_accumulate_forward_backward(dev):
    import nerfacc
    g = torch.Generator().manual_seed(0)
    n_rays = 500
    cnt = torch.randint(0, 40, (n_rays,), generator=g)
    ri = torch.repeat_interleave(torch.arange(n_rays), cnt)
    S = ri.numel()
    ts = torch.rand(S, generator=g) * 3; te = ts + 0.005 + torch.rand(S, generator=g) * 0.01
    sig = torch.rand(S, generator=g) * 50
    vals = torch.rand(S, 3, generator=g)
    gc = torch.rand(n_rays, 3, generator=g); ga = torch.rand(n_rays, 1, generator=g); gt = torch.rand(S, generator=g) * 0.1
    sr = sig.clone().requires_grad_(True); vr = vals.clone().requires_grad_(True)
    w, T, a = G.render_weight_from_density(ts, te, sr, ri, n_rays)
    col = G.accumulate_along_rays(w, vr, ri, n_rays); alp = G.accumulate_along_rays(w, None, ri, n_rays)
    ((col * gc).sum() + (alp * ga).sum() + (T * gt).sum()).backward()
    sg = sig.to(dev).requires_grad_(True); vg = vals.to(dev).requires_grad_(True)
    w2, T2, a2 = nerfacc.render_weight_from_density(ts.to(dev), te.to(dev), sg, ray_indices=ri.to(dev), n_rays=n_rays)
    col2 = nerfacc.accumulate_along_rays(w2, values=vg, ray_indices=ri.to(dev), n_rays=n_rays)
    alp2 = nerfacc.accumulate_along_rays(w2, values=None, ray_indices=ri.to(dev), n_rays=n_rays)
    ((col2 * gc.to(dev)).sum() + (alp2 * ga.to(dev)).sum() + (T2 * gt.to(dev)).sum()).backward()
    assert float((w2.detach().cpu() - w.detach()).abs().max()) < 1e-6 and float((T2.detach().cpu() - T.detach()).abs().max()) < 1e-6
    assert float((col2.detach().cpu() - col.detach()).abs().max()) < 1e-5 and float((alp2.detach().cpu() - alp.detach()).abs().max()) < 1e-5
    assert _rel(sg.grad, sr.grad) < 1e-4 and _rel(vg.grad, vr.grad) < 1e-5


def test_render_nerf_call_sequence_matches_oracle(dev):
    """InstantNGP.render_nerf (Instant_NGP.py:101-156) with the shims vs the same sequence on the oracle."""
    import nerfacc
    from kiui.gridencoder import GridEncoder
    from kiui.nn import MLP, trunc_exp
    torch.manual_seed(0)
    h = w = 32
    enc_d, enc_c = GridEncoder(num_levels=12).to(dev), GridEncoder(num_levels=12).to(dev)
    enc_d.embeddings.data.uniform_(-0.5, 0.5); enc_c.embeddings.data.uniform_(-0.5, 0.5)
    mlp_d, mlp_c = MLP(24, 1, 32, 2, bias=False).to(dev), MLP(24, 3, 32, 2, bias=False).to(dev)
    est = nerfacc.OccGridEstimator(roi_aabb=torch.tensor([-1.0, -1, -1, 1, 1, 1], device=dev), resolution=64, levels=1).to(dev)
    est.binaries = _binary_ball(64, 0.6).to(dev)[None]
    est.occs = est.binaries.flatten().float()
    ro, rd = G.get_rays(O.orbit_camera(0, 30, 1.75), h, w, 49.1)
    rog, rdg = ro.to(dev), rd.to(dev)

    def density(xs):
        return trunc_exp(mlp_d(enc_d(xs))).squeeze(-1)

    def sigma_fn(t0, t1, ri):
        return density(rog[ri] + rdg[ri] * (t0 + t1)[:, None] / 2.0)
    with torch.no_grad():
        ri, t0, t1 = est.sampling(rog, rdg, sigma_fn=sigma_fn, near_plane=0.01, far_plane=100, render_step_size=5e-3,
                                  stratified=False, cone_angle=0)
    xs = rog[ri] + rdg[ri] * (t0 + t1)[:, None] / 2.0
    sig = density(xs); rgb = torch.sigmoid(mlp_c(enc_c(xs)))
    wts, T, al = nerfacc.render_weight_from_density(t0, t1, sig, ray_indices=ri, n_rays=h * w)
    color = nerfacc.accumulate_along_rays(wts, values=rgb, ray_indices=ri, n_rays=h * w)
    alpha = nerfacc.accumulate_along_rays(wts, values=None, ray_indices=ri, n_rays=h * w)
    color = color + (1.0 - alpha) * 1.0
    gimg = torch.rand(h * w, 3, generator=torch.Generator().manual_seed(5))
    ((color * gimg.to(dev)).sum() + 0.1 * (alpha ** 2).sum()).backward()

    # oracle replay with the same parameters and the same sample list
    off = G.grid_offsets(num_levels=12)
    ed = enc_d.embeddings.detach().cpu().clone().requires_grad_(True); ec = enc_c.embeddings.detach().cpu().clone().requires_grad_(True)
    Wd = [l.weight.detach().cpu() for l in mlp_d.net]; Wc = [l.weight.detach().cpu() for l in mlp_c.net]
    def mlp(W, x):
        return torch.relu(x @ W[0].T) @ W[1].T
    ri_r, t0_r, t1_r = G.march(ro, rd, _binary_ball(64, 0.6), torch.tensor([-1.0, -1, -1, 1, 1, 1]), 0.01, 100.0, 5e-3)
    xs_r = ro[ri_r] + rd[ri_r] * (t0_r + t1_r)[:, None] / 2.0
    with torch.no_grad():
        s0 = G.trunc_exp(mlp(Wd, G.grid_encode(xs_r, ed.detach(), off, num_levels=12))).squeeze(-1)
        m = G.visibility_mask(t0_r, t1_r, s0, ri_r, h * w, 1e-4, 0.0)
    ri_r, t0_r, t1_r = ri_r[m], t0_r[m], t1_r[m]
    assert ri.numel() == ri_r.numel() and torch.equal(ri.cpu(), ri_r) and torch.equal(t0.cpu(), t0_r)
    xs_r = ro[ri_r] + rd[ri_r] * (t0_r + t1_r)[:, None] / 2.0
    sig_r = G.trunc_exp(mlp(Wd, G.grid_encode(xs_r, ed, off, num_levels=12))).squeeze(-1)
    rgb_r = torch.sigmoid(mlp(Wc, G.grid_encode(xs_r, ec, off, num_levels=12)))
    w_r, _, _ = G.render_weight_from_density(t0_r, t1_r, sig_r, ri_r, h * w)
    col_r = G.accumulate_along_rays(w_r, rgb_r, ri_r, h * w); alp_r = G.accumulate_along_rays(w_r, None, ri_r, h * w)
    col_r = col_r + (1.0 - alp_r)
    ((col_r * gimg).sum() + 0.1 * (alp_r ** 2).sum()).backward()
    assert float((color.detach().cpu() - col_r.detach()).abs().max()) < 1e-4
    assert float((alpha.detach().cpu() - alp_r.detach()).abs().max()) < 1e-4
    assert _rel(enc_d.embeddings.grad, ed.grad) < 1e-3 and _rel(enc_c.embeddings.grad, ec.grad) < 1e-3


def test_occupancy_grid_update_host_logic(dev):
    import nerfacc
    est = nerfacc.OccGridEstimator(roi_aabb=torch.tensor([-1.0, -1, -1, 1, 1, 1], device=dev), resolution=16, levels=1).to(dev)
    est.train()
    ball = lambda x: (x.norm(dim=-1, keepdim=True) < 0.5).float() * 0.05
    est.update_every_n_steps(0, occ_eval_fn=ball, occ_thre=0.01, n=8)
    inside = _binary_ball(16, 0.35).to(dev); outside = ~_binary_ball(16, 0.7).to(dev)
    assert bool(est.binaries[0][inside].all()) and not bool(est.binaries[0][outside].any())
    est.update_every_n_steps(3, occ_eval_fn=ball, occ_thre=0.01, n=8)        # not a multiple of n: no-op
    est.update_every_n_steps(512, occ_eval_fn=ball, occ_thre=0.01, n=8)      # post-warm-up branch
    assert bool(est.binaries[0][inside].all())


@pytest.mark.parametrize("din,dout,n", [(24, 1, 1000), (24, 3, 4097), (32, 1, 70000), (32, 3, 128), (32, 4, 333)])
def test_fused_mlp_matches_linear_stack(dev, din, dout, n):
    from kiui.nn import MLP
    torch.manual_seed(din + dout)
    m = MLP(din, dout, 32, 2, bias=False).to(dev)
    x = torch.randn(n, din, device=dev)
    xr = x.clone().requires_grad_(True); xf = x.clone().requires_grad_(True)
    ref = torch.relu(xr @ m.net[0].weight.T) @ m.net[1].weight.T
    out = m(xf)                                                   # fused path
    assert out.shape == (n, dout) and float((out - ref).abs().max()) < 1e-4 * max(1.0, float(ref.abs().max()))
    g = torch.randn(n, dout, device=dev)
    gw_ref = torch.autograd.grad((ref * g).sum(), [xr, m.net[0].weight, m.net[1].weight])
    gw = torch.autograd.grad((out * g).sum(), [xf, m.net[0].weight, m.net[1].weight])
    for a, b in zip(gw, gw_ref):
        assert float((a - b).norm() / (b.norm() + 1e-30)) < 1e-4
    # shapes the fused kernel does not cover fall back to the Linear stack
    m2 = MLP(16, 2, 64, 3, bias=True).to(dev)
    assert m2(torch.randn(10, 16, device=dev)).shape == (10, 2)
    assert m(x.view(10, -1, din)).shape[-1] == dout if n % 10 == 0 else True


def test_fit_nerf_steps_reproduce_the_reference_loop(dev):
    """End to end for path B: InstantNGP.fit_nerf (Instant_NGP.py:158-205) was executed from the reference source for two
    steps on the CPU with nerfacc / kiui served by the oracle (tests/golden/make_golden_training.py -> ref_ngp_fit.npz).
    The same steps through the shims (CUDA hash-grid encode / scatter / TV, marcher, volume rendering, fused MLPs) with
    the same random draws (gs_b200.ngp.set_rng) must land on the same parameters."""
    import os, random
    import nerfacc
    from conftest import GOLDEN
    from gs_b200 import ngp as NG
    from kiui.gridencoder import GridEncoder
    from kiui.nn import MLP, trunc_exp
    F_ = np.load(os.path.join(GOLDEN, "ref_ngp_fit.npz"))
    res, K, L = int(F_["res"]), int(F_["K"]), int(F_["levels"])
    gen = torch.Generator()                                   # CPU generator = the fixture's global CPU stream
    gen.manual_seed(int(F_["seed"]))
    enc_d, enc_c = GridEncoder(num_levels=L).to(dev), GridEncoder(num_levels=L).to(dev)
    init = []
    for enc in (enc_d, enc_c):
        e = torch.empty(enc.embeddings.shape).uniform_(-0.5, 0.5, generator=gen); init.append(e); enc.embeddings.data.copy_(e)
    mlp_d, mlp_c = MLP(2 * L, 1, 32, 2, bias=False).to(dev), MLP(2 * L, 3, 32, 2, bias=False).to(dev)
    for mlp in (mlp_d, mlp_c):
        for lin in mlp.net:
            lin.weight.data.copy_(torch.empty(lin.weight.shape).uniform_(-0.4, 0.4, generator=gen))
    # the fixture drew the reference images / masks next from the same stream
    n_ref = F_["ref_imgs"].shape[0]
    ref_imgs = [torch.rand(res, res, 3, generator=gen) for _ in range(n_ref)]
    ref_masks = [(torch.rand(res, res, generator=gen) > 0.4).float() for _ in range(n_ref)]
    assert np.array_equal(torch.stack(ref_imgs).numpy(), F_["ref_imgs"]) and np.array_equal(torch.stack(ref_masks).numpy(), F_["ref_masks"])
    img_gt = torch.stack(ref_imgs).permute(0, 3, 1, 2).contiguous().to(dev)       # prepare_torch_img at native size = identity resample
    msk_gt = torch.stack(ref_masks).to(dev)
    est = nerfacc.OccGridEstimator(roi_aabb=torch.tensor([-1.0, -1, -1, 1, 1, 1], device=dev), resolution=64, levels=1).to(dev)
    est.train()
    opt = torch.optim.Adam([{"params": enc_d.parameters(), "lr": 1e-2}, {"params": enc_c.parameters(), "lr": 1e-2},
                            {"params": mlp_d.parameters(), "lr": 1e-3}, {"params": mlp_c.parameters(), "lr": 1e-3}])
    density = lambda xs: trunc_exp(mlp_d(enc_d(xs)))
    NG.set_rng(gen)
    try:
        for step in range(K):
            i = int(F_["idx"][step])
            radius, elev, azim, cx, cy, cz = [float(v) for v in F_["poses"][i]]
            pose = O.orbit_camera(elev, azim, radius, target=np.array([cx, cy, cz], dtype=np.float32))
            ro, rd = G.get_rays(pose, res, res, 49.1)
            rog, rdg = ro.to(dev), rd.to(dev)
            est.update_every_n_steps(step, occ_eval_fn=lambda xs: density(xs) * 5e-3, occ_thre=0.01, n=8)

            def sigma_fn(t0, t1, ri):
                return density(rog[ri] + rdg[ri] * (t0 + t1)[:, None] / 2.0).squeeze(-1)
            with torch.no_grad():
                ri, t0, t1 = est.sampling(rog, rdg, sigma_fn=sigma_fn, near_plane=0.01, far_plane=100, render_step_size=5e-3,
                                          stratified=True, cone_angle=0)
            xs = rog[ri] + rdg[ri] * (t0 + t1)[:, None] / 2.0
            sig = density(xs).squeeze(-1); rgb = torch.sigmoid(mlp_c(enc_c(xs)))
            w, _, _ = nerfacc.render_weight_from_density(t0, t1, sig, ray_indices=ri, n_rays=res * res)
            color = nerfacc.accumulate_along_rays(w, values=rgb, ray_indices=ri, n_rays=res * res)
            alpha = nerfacc.accumulate_along_rays(w, values=None, ray_indices=ri, n_rays=res * res)
            color = color + (1.0 - alpha) * 1.0
            image_pred = color.view(res, res, 3).clamp(0, 1).permute(2, 0, 1).contiguous()
            alpha_pred = alpha.view(res, res).clamp(0, 1).contiguous()
            loss = torch.nn.functional.mse_loss(image_pred, img_gt[i]) + 0.1 * torch.nn.functional.mse_loss(alpha_pred, msk_gt[i])
            loss.backward()
            enc_d.grad_total_variation(1e-8)
            opt.step(); opt.zero_grad()
    finally:
        NG.set_rng(None)
    # the occupancy threshold is the MEAN density of 262k jittered cells: a cell sitting on it may flip between CPU and GPU
    assert abs(int(est.binaries.sum()) - int(F_["binaries_count"])) <= 0.005 * int(F_["binaries_count"])
    for name, mlp in (("mlp_density", mlp_d), ("mlp_color", mlp_c)):
        for l, lin in enumerate(mlp.net):
            d = np.abs(lin.weight.detach().cpu().numpy() - F_[f"{name}_w{l}"])
            assert float((d > 2e-4).mean()) <= 0.01 and float(d.max()) <= 2.0 * 1e-3 * K + 2e-4, (name, l, float(d.max()))
    for name, enc, e0 in (("emb_density", enc_d, init[0]), ("emb_color", enc_c, init[1])):
        emb = enc.embeddings.detach().cpu()
        ids = torch.from_numpy(F_[name + "_ids"])
        d = (emb[ids] - torch.from_numpy(F_[name + "_vals"])).abs().numpy()
        # Adam (lr 1e-2) turns gradient signs into +-lr steps: entries whose only gradient is the 1e-8 TV term, or summation
        # noise, can flip; everything else agrees to rounding
        assert float((d > 2e-4).mean()) <= 0.01 and float(d.max()) <= 2.0 * 1e-2 * K + 2e-4, (name, float(d.max()), float((d > 2e-4).mean()))
        n_changed = int((emb != e0).any(dim=1).sum())
        assert abs(n_changed - int(F_[name + "_n_changed"])) <= 0.01 * int(F_[name + "_n_changed"])
        assert abs(float(emb.double().sum()) - float(F_[name + "_sum"])) <= 1e-6 * emb.numel() * 0.5 + 2e-2 * 0.01 * n_changed
