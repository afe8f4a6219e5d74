"""Multi-view optimisation step around the rasterizer (the metric's timed region).

What GaussianSplatting3D.training does per step for its `batch_size` views
(MVs_Algorithms/GaussianSplatting/main_3DGS.py:158-205): render each view, backprop,
sum the gradients — here with the summation done in place on the device
(`accumulate=1` of gs_b200_rasterize_backward) into ONE packed gradient buffer laid out
means3D | shs | opacities | scales | rotations | means2D, which is also the buffer the
data-parallel all-reduce uses.
"""
import ctypes as C
import torch

from . import _lib
from .rasterizer import _Buffers, _ptr, _stream


class PackedParams:
    """Activated Gaussian parameters + packed gradient buffer on one device."""

    def __init__(self, cloud: dict):
        self.means3D = cloud["means3D"].contiguous().float()
        self.shs = cloud["shs"].contiguous().float()
        self.opacities = cloud["opacities"].contiguous().float()
        self.scales = cloud["scales"].contiguous().float()
        self.rotations = cloud["rotations"].contiguous().float()
        self.N = self.means3D.shape[0]
        self.M = self.shs.shape[1]
        dev = self.means3D.device
        n = self.N
        sizes = [3 * n, 3 * self.M * n, n, 3 * n, 4 * n, 3 * n]
        self.grads = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        o = 0
        views = []
        for s in sizes:
            views.append(self.grads[o:o + s]); o += s
        self.g_means3D, self.g_shs, self.g_opacities, self.g_scales, self.g_rotations, self.g_means2D = views
        self.radii = torch.empty(n, dtype=torch.int32, device=dev)

    def param_bytes(self):
        return 4 * self.N * (3 + 3 * self.M + 1 + 3 + 4)

    def grad_bytes(self):
        return 4 * self.grads.numel()


class ViewSet:
    """V views on the device: packed [V,40] records (camera.orbit_views) + image size."""

    def __init__(self, views_np, W, H, sh_degree, device, scale_modifier=1.0):
        self.host = views_np
        self.dev = torch.from_numpy(views_np).to(device).contiguous()
        self.W, self.H, self.V = W, H, views_np.shape[0]
        self.sh_degree, self.scale_modifier = sh_degree, scale_modifier

    def view(self, v) -> _lib.View:
        base = self.dev.data_ptr() + v * 160
        return _lib.View(self.H, self.W, float(self.host[v, 38]), float(self.host[v, 39]), base + 35 * 4,
                         float(self.scale_modifier), base, base + 16 * 4, self.sh_degree, base + 32 * 4, 0, 0)


def step_device(params: PackedParams, views: ViewSet, dL_dout: torch.Tensor, out_images: torch.Tensor = None):
    """fwd+bwd over all views with device-resident inputs; gradients summed into params.grads.

    dL_dout: [V,5,H,W] (3 colour planes, depth, alpha).  Returns total (tile,splat) pairs.
    """
    dev = params.means3D.device
    H, W, V = views.H, views.W, views.V
    npix = H * W
    params.grads.zero_()
    if out_images is None:
        out_images = torch.empty(5, H, W, dtype=torch.float32, device=dev)
        per_view = False
    else:
        per_view = True
    total_pairs = 0
    stream = _stream()
    for v in range(V):
        view = views.view(v)
        img = out_images[v] if per_view else out_images
        ip = img.data_ptr()
        bufs = _Buffers(dev)
        state = _lib.State()
        rc = _lib.lib.gs_b200_rasterize_forward(
            C.byref(view), params.N, params.M, _ptr(params.means3D), _ptr(params.shs), None, _ptr(params.opacities),
            _ptr(params.scales), _ptr(params.rotations), None, C.c_void_p(ip), C.c_void_p(ip + 12 * npix),
            C.c_void_p(ip + 16 * npix), _ptr(params.radii), bufs.cb, None, C.byref(state), stream)
        bufs.scratch.clear()
        _lib.check(rc)
        total_pairs += int(state.num_rendered)
        up = dL_dout[v].data_ptr()
        rc = _lib.lib.gs_b200_rasterize_backward(
            C.byref(view), params.N, params.M, _ptr(params.means3D), _ptr(params.shs), None, _ptr(params.opacities),
            _ptr(params.scales), _ptr(params.rotations), None, _ptr(params.radii), C.byref(state),
            C.c_void_p(up), C.c_void_p(up + 12 * npix), C.c_void_p(up + 16 * npix),
            _ptr(params.g_means3D), _ptr(params.g_means2D), _ptr(params.g_shs), None, _ptr(params.g_opacities),
            _ptr(params.g_scales), _ptr(params.g_rotations), None, 1, bufs.cb, None, stream)
        bufs.scratch.clear()
        _lib.check(rc)
    return total_pairs


def step_device_pipelined(params: PackedParams, views: ViewSet, dL_dout: torch.Tensor, out_images: torch.Tensor = None):
    """Same contract as step_device, through the C entry gs_b200_step_device: views software-pipelined over two
    internal streams with persistent workspaces (no per-view allocation, host pair-count wait hidden)."""
    import numpy as np
    assert views.host.dtype == np.float32 and views.host.flags["C_CONTIGUOUS"]
    pairs = C.c_int64(0)
    with torch.cuda.device(params.means3D.device):        # the pipeline state is per device (streams, workspaces)
        _lib.check(_lib.lib.gs_b200_step_device(
            views.V, views.H, views.W, views.sh_degree, float(views.scale_modifier), C.c_void_p(views.host.ctypes.data),
            _ptr(views.dev), params.N, params.M, _ptr(params.means3D), _ptr(params.shs), _ptr(params.opacities),
            _ptr(params.scales), _ptr(params.rotations), _ptr(dL_dout), _ptr(params.grads),
            None if out_images is None else _ptr(out_images), C.byref(pairs), _stream()))
    return int(pairs.value)


def step_device_loss(params: PackedParams, views: ViewSet, loss_grad_fn, images: torch.Tensor, dL_dout: torch.Tensor,
                     radii: torch.Tensor = None):
    """Forward + loss + backward in ONE pipelined pass (gs_b200_step_device_hook).

    loss_grad_fn(v, image[5,H,W], dL[5,H,W]) is called once per view, with torch's current stream set to the internal
    stream that just rendered images[v]; it must fill dL (in place) with d loss / d image.  Everything it enqueues
    runs between that view's forward and backward, overlapped with the other views' kernels.  Returns pair count."""
    import numpy as np
    assert views.host.dtype == np.float32 and views.host.flags["C_CONTIGUOUS"]
    assert images.shape == dL_dout.shape == (views.V, 5, views.H, views.W) and images.is_contiguous() and dL_dout.is_contiguous()
    if radii is not None:
        assert radii.dtype == torch.int32 and radii.shape == (views.V, params.N) and radii.is_contiguous()
    err = []

    def _hook(_user, v, stream_ptr):
        try:
            with torch.cuda.stream(torch.cuda.ExternalStream(int(stream_ptr), device=images.device)):
                loss_grad_fn(int(v), images[v], dL_dout[v])
            return 0
        except BaseException as e:     # never let an exception cross the C frame
            err.append(e)
            return 1
    cb = _lib.VIEW_HOOK(_hook)
    pairs = C.c_int64(0)
    with torch.cuda.device(params.means3D.device):
        rc = _lib.lib.gs_b200_step_device_hook(
            views.V, views.H, views.W, views.sh_degree, float(views.scale_modifier), C.c_void_p(views.host.ctypes.data),
            _ptr(views.dev), params.N, params.M, _ptr(params.means3D), _ptr(params.shs), _ptr(params.opacities),
            _ptr(params.scales), _ptr(params.rotations), _ptr(dL_dout), _ptr(params.grads), _ptr(images),
            None if radii is None else _ptr(radii), cb, None, C.byref(pairs), _stream())
    if err:
        raise err[0]
    _lib.check(rc)
    return int(pairs.value)


def image_loss(image: torch.Tensor, ref_image: torch.Tensor, ref_mask: torch.Tensor, lambda_ssim=0.2, lambda_alpha=3.0, scale=1.0):
    """CUDA training loss of one view (gs_b200_image_loss): image [5,H,W] -> (loss 0-d tensor, dL/dimage [5,H,W])."""
    _, H, W = image.shape
    assert image.shape[0] == 5 and ref_image.shape == (3, H, W) and ref_mask.shape == (1, H, W)
    for t in (image, ref_image, ref_mask):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
    dl = torch.empty_like(image); loss = torch.empty(1, device=image.device)
    _lib.check(_lib.lib.gs_b200_image_loss(H, W, _ptr(image), _ptr(ref_image), _ptr(ref_mask), float(lambda_ssim), float(lambda_alpha),
                                           float(scale), _ptr(dl), _ptr(loss), _stream()))
    return loss[0], dl


def step_device_train(params: PackedParams, views: ViewSet, ref_images: torch.Tensor, ref_masks: torch.Tensor, lambda_ssim,
                      lambda_alpha, loss_scale, images: torch.Tensor, dL_dout: torch.Tensor, losses: torch.Tensor,
                      radii: torch.Tensor = None):
    """forward -> CUDA loss + gradient -> backward per view inside the pipeline (gs_b200_step_device_train)."""
    import numpy as np
    V, H, W = views.V, views.H, views.W
    assert views.host.dtype == np.float32 and views.host.flags["C_CONTIGUOUS"]
    assert images.shape == dL_dout.shape == (V, 5, H, W) and ref_images.shape == (V, 3, H, W) and ref_masks.shape == (V, 1, H, W)
    assert losses.shape == (V,) and losses.dtype == torch.float32
    for t in (images, dL_dout, ref_images, ref_masks, losses):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
    if radii is not None:
        assert radii.dtype == torch.int32 and radii.shape == (V, params.N) and radii.is_contiguous()
    pairs = C.c_int64(0)
    with torch.cuda.device(params.means3D.device):
      _lib.check(_lib.lib.gs_b200_step_device_train(
        V, H, W, views.sh_degree, float(views.scale_modifier), C.c_void_p(views.host.ctypes.data), _ptr(views.dev),
        params.N, params.M, _ptr(params.means3D), _ptr(params.shs), _ptr(params.opacities), _ptr(params.scales),
        _ptr(params.rotations), _ptr(ref_images), _ptr(ref_masks), float(lambda_ssim), float(lambda_alpha), float(loss_scale),
        _ptr(dL_dout), _ptr(params.grads), _ptr(images), None if radii is None else _ptr(radii), _ptr(losses),
        C.byref(pairs), _stream()))
    return int(pairs.value)


def render_views(params: PackedParams, views: ViewSet, images: torch.Tensor = None, radii: torch.Tensor = None,
                 colors_precomp: torch.Tensor = None):
    """Forward only over all views (gs_b200_render_views).  Returns (images[V,5,H,W], pair count); `radii`
    (optional int32 [V,N]) receives the per-view radii (visibility filter = radii > 0).  colors_precomp [N,3]
    replaces params.shs (LGM-style callers, Gen_3D_Modules/LGM/core/gs.py:75-84)."""
    import numpy as np
    assert views.host.dtype == np.float32 and views.host.flags["C_CONTIGUOUS"]
    if images is None:
        images = torch.empty(views.V, 5, views.H, views.W, dtype=torch.float32, device=params.means3D.device)
    assert images.shape == (views.V, 5, views.H, views.W) and images.is_contiguous()
    if radii is not None:
        assert radii.dtype == torch.int32 and radii.shape == (views.V, params.N) and radii.is_contiguous()
    pairs = C.c_int64(0)
    with torch.cuda.device(params.means3D.device):
      _lib.check(_lib.lib.gs_b200_render_views(
        views.V, views.H, views.W, views.sh_degree, float(views.scale_modifier), C.c_void_p(views.host.ctypes.data),
        _ptr(views.dev), params.N, params.M, _ptr(params.means3D), None if colors_precomp is not None else _ptr(params.shs),
        None if colors_precomp is None else _ptr(colors_precomp.contiguous().float()), _ptr(params.opacities),
        _ptr(params.scales), _ptr(params.rotations), _ptr(images), None if radii is None else _ptr(radii),
        C.byref(pairs), _stream()))
    return images, int(pairs.value)


class HostStep:
    """The e2e entry (gs_b200_step_host): pinned host buffers in, summed gradients out."""

    def __init__(self, cloud_cpu: dict, views_np, W, H, sh_degree, dL_dout_cpu: torch.Tensor, scale_modifier=1.0):
        pin = lambda t: t.contiguous().float().pin_memory()
        self.means3D, self.shs = pin(cloud_cpu["means3D"]), pin(cloud_cpu["shs"])
        self.opacities, self.scales, self.rotations = pin(cloud_cpu["opacities"]), pin(cloud_cpu["scales"]), pin(cloud_cpu["rotations"])
        self.views = torch.from_numpy(views_np).contiguous().pin_memory()
        self.dL = pin(dL_dout_cpu)
        self.N, self.M = self.means3D.shape[0], self.shs.shape[1]
        self.V, self.W, self.H, self.sh_degree, self.scale_modifier = views_np.shape[0], W, H, sh_degree, scale_modifier
        self.n_grad = self.N * (3 + 3 * self.M + 1 + 3 + 4) + 3 * self.N
        self.grads = torch.empty(self.n_grad, dtype=torch.float32).pin_memory()
        self.h2d_bytes = 4 * (self.N * (3 + 3 * self.M + 1 + 3 + 4) + self.views.numel() + self.dL.numel())
        self.d2h_bytes = 4 * self.n_grad

    def _args(self):
        return (self.V, self.H, self.W, self.sh_degree, float(self.scale_modifier), _ptr(self.views), self.N, self.M,
                _ptr(self.means3D), _ptr(self.shs), _ptr(self.opacities), _ptr(self.scales), _ptr(self.rotations),
                _ptr(self.dL))

    def run(self, images: torch.Tensor = None):
        """images (optional): pinned host [V,5,H,W] fp32 that receives the rendered colour | depth | alpha planes."""
        pairs = C.c_int64(0)
        _lib.check(_lib.lib.gs_b200_step_host(*self._args(), _ptr(self.grads), _ptr(images) if images is not None else None,
                                              C.byref(pairs), _stream()))
        return int(pairs.value)

    def run_dev_grads(self, grads_dev: torch.Tensor):
        pairs = C.c_int64(0)
        _lib.check(_lib.lib.gs_b200_step_host_dev_grads(*self._args(), _ptr(grads_dev), None, C.byref(pairs), _stream()))
        return int(pairs.value)
