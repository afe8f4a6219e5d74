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
