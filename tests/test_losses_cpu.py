"""CPU tests of the image losses (torch): MS-SSIM restatement properties and the training loss composition."""
import pytest
import torch

from conftest import ROOT  # noqa: F401
from gs_b200 import losses


def test_ms_ssim_properties():
    g = torch.Generator().manual_seed(0)
    x = torch.rand(2, 3, 176, 200, generator=g)
    assert abs(float(losses.ms_ssim(x, x)) - 1.0) < 1e-6
    y = (x + 0.1 * torch.randn(x.shape, generator=g)).clamp(0, 1)
    z = (x + 0.3 * torch.randn(x.shape, generator=g)).clamp(0, 1)
    sxy, syx, sxz = float(losses.ms_ssim(x, y)), float(losses.ms_ssim(y, x)), float(losses.ms_ssim(x, z))
    assert abs(sxy - syx) < 1e-6 and 0 < sxz < sxy < 1
    with pytest.raises(ValueError):
        losses.ms_ssim(x[..., :100, :100], y[..., :100, :100])


def test_single_scale_ssim_constant_images():
    # two constant images a, b: luminance term only -> (2ab + C1)/(a^2 + b^2 + C1), cs = 1
    a, b = 0.3, 0.6
    X = torch.full((1, 1, 32, 32), a); Y = torch.full((1, 1, 32, 32), b)
    ssim, cs = losses._ssim_cs(X, Y, losses._gauss_1d())
    C1 = 0.01 ** 2
    assert abs(float(ssim) - (2 * a * b + C1) / (a * a + b * b + C1)) < 1e-5 and abs(float(cs) - 1.0) < 1e-5


def test_training_loss_composition_and_gradient():
    g = torch.Generator().manual_seed(1)
    img = torch.rand(1, 3, 176, 176, generator=g, requires_grad=True)
    alpha = torch.rand(1, 1, 176, 176, generator=g, requires_grad=True)
    ref = torch.rand(1, 3, 176, 176, generator=g); mask = (torch.rand(1, 1, 176, 176, generator=g) > 0.5).float()
    l = losses.training_loss(img, alpha, ref, mask, lambda_ssim=0.2, lambda_alpha=3.0)
    expect = 0.8 * (img * mask - ref * mask).abs().mean() + 3.0 * ((alpha - mask) ** 2).mean() + 0.2 * (1 - losses.ms_ssim(ref * mask, img * mask))
    assert abs(float(l) - float(expect)) < 1e-6
    l.backward()
    assert torch.isfinite(img.grad).all() and float(img.grad.abs().max()) > 0 and float(alpha.grad.abs().max()) > 0
