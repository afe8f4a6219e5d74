"""GS .ply wire format: the on-disk contract between the reference's 3DGS nodes —
`GaussianModel.to_ply` / `create_from_ply` (main_3DGS_renderer.py:475-498) and `write_gs_ply` / `read_gs_ply`
(mesh_processer/mesh_utils.py:333-390): one `vertex` element, all properties float32, RAW (pre-activation) values, in
the order  x y z nx ny nz f_dc_0..2 f_rest_0..(3(M-1)-1) opacity scale_0..2 rot_0..3 ; SH planes are channel-major
(f_rest_{c*(M-1)+k} = coefficient k+1 of channel c).  Binary little-endian, written/read with numpy structured
arrays straight from/to the packed device buffers (no per-row Python tuples, no plyfile dependency)."""
import numpy as np
import torch


def attribute_names(M: int):
    names = ["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(3)] + [f"f_rest_{i}" for i in range(3 * (M - 1))]
    return names + ["opacity"] + [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)]


def pack_rows(xyz, shs, opacity, scaling, rotation) -> torch.Tensor:
    """[N, 14+3M] float32 rows in file order, built on the tensors' device. shs: [N,M,3] (coefficient-major rows)."""
    N, M = shs.shape[0], shs.shape[1]
    f_dc = shs[:, :1, :].transpose(1, 2).reshape(N, 3)
    f_rest = shs[:, 1:, :].transpose(1, 2).reshape(N, 3 * (M - 1))
    return torch.cat([xyz, torch.zeros_like(xyz), f_dc, f_rest, opacity.reshape(N, 1), scaling, rotation], dim=1).float().contiguous()


def unpack_rows(rows: torch.Tensor):
    N, K = rows.shape
    M = (K - 14) // 3
    xyz = rows[:, 0:3]
    f_dc = rows[:, 6:9].reshape(N, 3, 1)
    f_rest = rows[:, 9:9 + 3 * (M - 1)].reshape(N, 3, M - 1)
    shs = torch.cat([f_dc, f_rest], dim=2).transpose(1, 2).contiguous()            # [N,M,3]
    o = 9 + 3 * (M - 1)
    return dict(xyz=xyz.contiguous(), shs=shs, opacity=rows[:, o:o + 1].contiguous(), scaling=rows[:, o + 1:o + 4].contiguous(),
                rotation=rows[:, o + 4:o + 8].contiguous())


def write_gs_ply(path, xyz, shs, opacity, scaling, rotation):
    rows = pack_rows(xyz, shs, opacity, scaling, rotation).cpu().numpy().astype("<f4")
    names = attribute_names(shs.shape[1])
    assert rows.shape[1] == len(names)
    header = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % rows.shape[0]
    header += "".join(f"property float {n}\n" for n in names) + "end_header\n"
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(rows.tobytes())


def read_gs_ply(path, device="cpu"):
    with open(path, "rb") as f:
        data = f.read()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    lines = data[:end].decode("ascii").splitlines()
    if "format binary_little_endian 1.0" not in lines:
        raise ValueError("only binary_little_endian GS .ply files are supported")
    n = int([l for l in lines if l.startswith("element vertex")][0].split()[-1])
    props = [l.split()[-1] for l in lines if l.startswith("property float")]
    if len(props) != len([l for l in lines if l.startswith("property")]):
        raise ValueError("non-float properties are not part of the GS .ply contract")
    M = (len([p for p in props if p.startswith("f_rest_")]) + 3) // 3
    if props != attribute_names(M):
        raise ValueError("unexpected property order for a GS .ply")
    rows = np.frombuffer(data, dtype="<f4", count=n * len(props), offset=end).reshape(n, len(props))
    return unpack_rows(torch.from_numpy(rows.copy()).to(device))


def max_sh_degree_from_properties(n_rest: int) -> int:
    """calculate_max_sh_degree_from_gs_ply (mesh_utils.py:346-350)."""
    return int(((n_rest + 3) / 3) ** 0.5 - 1)


# ---------------------------------------------------------------------------------------------------------
# axis switch / rescale of a Gaussian cloud, on the device the tensors live on
# ---------------------------------------------------------------------------------------------------------
def quaternion_to_axis_angle(q: torch.Tensor) -> torch.Tensor:
    """(w, x, y, z) -> axis * angle, the convention of kornia.geometry.conversions.quaternion_to_axis_angle (kornia >= 0.7,
    the un-vendored package mesh_utils.py:3-6 imports; restated): angle in (-pi, pi], 2 * xyz for a zero vector part."""
    w, xyz = q[..., 0], q[..., 1:]
    s2 = (xyz * xyz).sum(-1)
    s = torch.sqrt(s2)
    two_theta = 2.0 * torch.where(w < 0, torch.atan2(-s, -w), torch.atan2(s, w))
    k = torch.where(s2 > 0, two_theta / s, torch.full_like(s, 2.0))
    return xyz * k[..., None]


def axis_angle_to_quaternion(a: torch.Tensor) -> torch.Tensor:
    """axis * angle -> (w, x, y, z), as kornia.geometry.conversions.axis_angle_to_quaternion (restated)."""
    t2 = (a * a).sum(-1)
    t = torch.sqrt(t2)
    half = 0.5 * t
    pos = t2 > 0
    k = torch.where(pos, torch.sin(half) / t, torch.full_like(t, 0.5))
    w = torch.where(pos, torch.cos(half), torch.ones_like(t))
    return torch.cat([w[..., None], a * k[..., None]], dim=-1)


def switch_axis_and_scale(fields: dict, target_axis, target_scale, coordinate_invert_count: int) -> dict:
    """switch_ply_axis_and_scale (mesh_processer/mesh_utils.py:446-472) without its numpy <-> CUDA round trips: every
    tensor stays on the device it is on.  `fields` = dict(xyz [N,3], shs [N,M,3], opacity [N,1], scaling [N,3] (raw log
    scales), rotation [N,4] (raw w,x,y,z)) as `read_gs_ply` / `unpack_rows` return; returns a new dict.
      xyz      <- (xyz * target_scale)[:, target_axis]
      scaling  <- scaling[:, target_axis]                      (the reference permutes the raw log-scales, no rescale)
      rotation <- quaternion of (axis_angle * target_scale)[:, target_axis], negated when the handedness flips an odd
                  number of times; SHs and opacity are untouched."""
    dev = fields["xyz"].device
    axis = torch.as_tensor(list(target_axis), dtype=torch.long, device=dev)
    scale = torch.as_tensor(list(target_scale), dtype=torch.float32, device=dev)
    out = dict(fields)
    out["xyz"] = (fields["xyz"].float() * scale)[:, axis].contiguous()
    out["scaling"] = fields["scaling"].float()[:, axis].contiguous()
    aa = (quaternion_to_axis_angle(fields["rotation"].float()) * scale)[:, axis]
    if coordinate_invert_count % 2 != 0:
        aa = -aa
    out["rotation"] = axis_angle_to_quaternion(aa).contiguous()
    return out
