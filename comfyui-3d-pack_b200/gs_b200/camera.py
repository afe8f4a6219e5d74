"""Host-side camera helpers mirroring the reference's conventions.

  orbit_camera ........ kiui.cam.orbit_camera as used at shared_utils/camera_utils.py:240-244
                        (same look-at construction as camera_utils.py:45-62)
  projection_matrix ... get_projection_matrix, camera_utils.py:174-185
  MiniCam ............. camera_utils.py:188-214 (w2c rectification, transposes, camera_center sign)
"""
import math
import numpy as np
import torch


def orbit_camera(elevation, azimuth, radius=1.0, is_degree=True, target=None, opengl=True):
    if is_degree:
        elevation = np.deg2rad(elevation); azimuth = np.deg2rad(azimuth)
    x = radius * np.cos(elevation) * np.sin(azimuth)
    y = -radius * np.sin(elevation)
    z = radius * np.cos(elevation) * np.cos(azimuth)
    if target is None:
        target = np.zeros([3], dtype=np.float32)
    campos = np.array([x, y, z], dtype=np.float32) + target
    def nrm(v):
        return v / max(float(np.linalg.norm(v)), 1e-20)
    if opengl:
        fwd = nrm(campos - target)
        up = np.array([0, 1, 0], dtype=np.float32)
        right = nrm(np.cross(up, fwd)); up = nrm(np.cross(fwd, right))
    else:
        fwd = nrm(target - campos)
        up = np.array([0, 1, 0], dtype=np.float32)
        right = nrm(np.cross(fwd, up)); up = nrm(np.cross(right, fwd))
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = np.stack([right, up, fwd], axis=1)
    T[:3, 3] = campos
    return T


def projection_matrix(znear, zfar, fovX, fovY, z_sign=1.0):
    P = torch.zeros(4, 4)
    P[0, 0] = 1 / math.tan(fovX / 2)
    P[1, 1] = 1 / math.tan(fovY / 2)
    P[3, 2] = z_sign
    P[2, 2] = z_sign * zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


class MiniCam:
    """Same attributes as the reference's MiniCam, tensors on `device`."""

    def __init__(self, c2w, width, height, fovy, fovx, znear, zfar, projection_matrix_t=None, device="cuda"):
        self.image_width, self.image_height = width, height
        self.FoVy, self.FoVx, self.znear, self.zfar = fovy, fovx, znear, zfar
        w2c = np.linalg.inv(np.asarray(c2w, dtype=np.float32))
        w2c[1:3, :3] *= -1
        w2c[:3, 3] *= -1
        self.world_view_transform = torch.tensor(w2c).transpose(0, 1).contiguous().to(device)
        self.projection_matrix = (projection_matrix(znear, zfar, fovx, fovy).transpose(0, 1).to(device)
                                  if projection_matrix_t is None else projection_matrix_t)
        self.full_proj_transform = (self.world_view_transform @ self.projection_matrix).contiguous()
        self.camera_center = -torch.tensor(np.asarray(c2w, dtype=np.float32)[:3, 3]).to(device)


def orbit_views(n_views, W, H, fovy_deg=49.1, radius=1.75, elevation=0.0, znear=0.01, zfar=100.0, bg=(0., 0., 0.),
                azimuth_offset=0.0, n_total=None, start=0):
    """Packed [n_views,40] fp32 view records (include/gs_b200.h, gs_b200_step_host) for an orbit ring:
    viewmatrix(16) projmatrix(16) campos(3) bg(3) tanfovx tanfovy; azimuth = 360*(start+k)/n_total."""
    n_total = n_views if n_total is None else n_total
    fovy = np.deg2rad(fovy_deg)
    fovx = 2 * np.arctan(np.tan(fovy / 2) * W / H)
    out = np.zeros((n_views, 40), dtype=np.float32)
    for k in range(n_views):
        az = azimuth_offset + 360.0 * (start + k) / n_total
        cam = MiniCam(orbit_camera(elevation, az, radius), W, H, fovy, fovx, znear, zfar, device="cpu")
        out[k, 0:16] = cam.world_view_transform.reshape(-1).numpy()
        out[k, 16:32] = cam.full_proj_transform.reshape(-1).numpy()
        out[k, 32:35] = cam.camera_center.numpy()
        out[k, 35:38] = bg
        out[k, 38] = math.tan(fovx * 0.5)
        out[k, 39] = math.tan(fovy * 0.5)
    return out
