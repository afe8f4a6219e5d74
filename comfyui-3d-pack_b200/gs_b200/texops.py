"""Texture-space helpers beside the mesh path.

`uv_padding` restates kiui.op.uv_padding (kiui 0.2.14, backend 'knn') as used by color_func_to_albedo
(mesh_processer/mesh_utils.py:521-568): texels outside the valid mask but within `padding` 4-connected dilation steps
of it (city-block distance <= padding) take the colour of the Euclidean-nearest valid texel; everything else is left
untouched.  The package does this on the CPU with scipy dilation + a kd-tree; here small paddings (the call site uses
2) run as (2p+1)^2 shifted-mask passes on whatever device the image lives on, large ones through a kd-tree.
"""
import numpy as np
import torch

_WINDOW_MAX = 8


def _offsets(p):
    offs = [(dy, dx) for dy in range(-p, p + 1) for dx in range(-p, p + 1) if (dy or dx) and dy * dy + dx * dx <= p * p]
    offs.sort(key=lambda o: (o[0] * o[0] + o[1] * o[1], o[0], o[1]))          # nearest first; ties: smallest (dy, dx)
    return offs


def _shift(t, dy, dx, fill):
    """out[y, x] = t[y + dy, x + dx], `fill` outside."""
    H, W = t.shape[:2]
    out = torch.full_like(t, fill)
    ys, ye = max(0, -dy), min(H, H - dy)
    xs, xe = max(0, -dx), min(W, W - dx)
    if ys < ye and xs < xe:
        out[ys:ye, xs:xe] = t[ys + dy:ye + dy, xs + dx:xe + dx]
    return out


def uv_padding(image, mask, padding=None, backend="knn"):
    """image [H,W,C] float in [0,1] (tensor or ndarray), mask [H,W] bool = valid texels -> padded image, same type."""
    is_t = torch.is_tensor(image)
    img = image if is_t else torch.from_numpy(np.asarray(image))
    m = (mask if torch.is_tensor(mask) else torch.from_numpy(np.asarray(mask))).to(img.device).bool()
    H, W = img.shape[:2]
    if padding is None:
        padding = int(0.1 * max(H, W))
    p = int(padding)
    if p <= 0 or not bool(m.any()) or bool(m.all()):
        out = img.clone()
    elif p <= _WINDOW_MAX:
        # region to fill: city-block distance <= p from the mask (p steps of 4-connected dilation)
        region = m.clone()
        for _ in range(p):
            region = region | _shift(region, 1, 0, False) | _shift(region, -1, 0, False) | _shift(region, 0, 1, False) | _shift(region, 0, -1, False)
        todo = region & ~m
        out = img.clone()
        for dy, dx in _offsets(p):
            if not bool(todo.any()):
                break
            hit = todo & _shift(m, dy, dx, False)                 # nearest valid texel of (y,x) is (y+dy, x+dx)
            if bool(hit.any()):
                out[hit] = _shift(img, dy, dx, 0.0)[hit]
                todo = todo & ~hit
    else:
        from scipy.ndimage import binary_dilation, binary_erosion
        from scipy.spatial import cKDTree
        mn = m.cpu().numpy()
        region = binary_dilation(mn, iterations=p) & ~mn
        band = mn & ~binary_erosion(mn, iterations=2)
        src = np.stack(np.nonzero(band), axis=-1)
        dst = np.stack(np.nonzero(region), axis=-1)
        outn = img.detach().cpu().numpy().copy()
        if len(dst):
            _, idx = cKDTree(src).query(dst, k=1)
            outn[tuple(dst.T)] = outn[tuple(src[idx].T)]
        out = torch.from_numpy(outn).to(img)
    return out if is_t else out.cpu().numpy()
