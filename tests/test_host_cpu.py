"""CPU tests of the host side (no GPU): the C-ABI library loads and exports every symbol the header declares,
struct layouts agree with the header, the reference-facing Python interface validates like the package does,
and the host camera logic matches the oracle / reference fixtures."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT

HEADER = os.path.join(ROOT, "include", "gs_b200.h")


def _ensure_built():
    lib = os.path.join(ROOT, "comfyui-3d-pack_b200", "gs_b200", "libgs_b200.so")
    if not os.path.exists(lib):
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()
    return lib


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_ensure_built())
    src = open(HEADER).read()
    names = set(re.findall(r"\b(gs_b200_[a-z0-9_]+)\s*\(", src))
    names -= {"gs_b200_alloc_fn"}
    assert len(names) >= 14
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/gs_b200.h but not exported"
    lib.gs_b200_abi_version.restype = ctypes.c_int32
    assert lib.gs_b200_abi_version() == 1


def test_every_header_is_plain_c_and_every_declared_entry_point_is_exported():
    """include/*.h (gs, dr, ngp): each compiles as C on its own (no C++ / torch types in the boundary) and every
    function it declares is an exported symbol of libgs_b200.so."""
    lib = ctypes.CDLL(_ensure_built())
    inc = os.path.join(ROOT, "include")
    headers = sorted(f for f in os.listdir(inc) if f.endswith(".h"))
    assert headers == ["dr_b200.h", "gs_b200.h", "ngp_b200.h"]
    total = 0
    for h in headers:
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", os.path.join(inc, h)], check=True)
        src = re.sub(r"/\*.*?\*/", "", open(os.path.join(inc, h)).read(), flags=re.S)
        prefix = h.split("_")[0]
        names = set(re.findall(r"\b(%s_b200_[a-z0-9_]+)\s*\(" % prefix, src))
        names -= {"gs_b200_alloc_fn"}
        assert names, h
        for n in sorted(names):
            assert hasattr(lib, n), f"{n} declared in include/{h} but not exported"
        total += len(names)
    assert total >= 45


def test_ctypes_structs_match_header_layout():
    from gs_b200 import _lib
    code = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "gs_b200.h"
    int main(void){ printf("%zu %zu %zu %zu %zu %zu\n", sizeof(gs_b200_view), offsetof(gs_b200_view, campos),
        sizeof(gs_b200_state), offsetof(gs_b200_state, num_rendered), offsetof(gs_b200_state, owned), offsetof(gs_b200_view, viewmatrix)); return 0; }
    '''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c"); exe = os.path.join(d, "t")
        open(c, "w").write(code)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    sv, oc, ss, onr, oo, ovm = map(int, out)
    assert ctypes.sizeof(_lib.View) == sv and _lib.View.campos.offset == oc and _lib.View.viewmatrix.offset == ovm
    assert ctypes.sizeof(_lib.State) == ss and _lib.State.num_rendered.offset == onr and _lib.State.owned.offset == oo


def test_settings_namedtuple_field_order_matches_reference_call_site():
    # main_3DGS_renderer.py:849-862 constructs it by keyword; LGM/TRELLIS too. Order = the package's NamedTuple.
    from diff_gaussian_rasterization import GaussianRasterizationSettings
    assert GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
        "sh_degree", "campos", "prefiltered", "debug")


def test_rasterizer_argument_validation_matches_package_messages():
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    z = torch.zeros
    rs = GaussianRasterizationSettings(16, 16, 1.0, 1.0, z(3), 1.0, torch.eye(4), torch.eye(4), 0, z(3), False, False)
    r = GaussianRasterizer(raster_settings=rs)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(means3D=z(1, 3), means2D=z(1, 3), opacities=z(1, 1), scales=z(1, 3), rotations=z(1, 4))
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(means3D=z(1, 3), means2D=z(1, 3), opacities=z(1, 1), shs=z(1, 1, 3), colors_precomp=z(1, 3), scales=z(1, 3), rotations=z(1, 4))
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(means3D=z(1, 3), means2D=z(1, 3), opacities=z(1, 1), colors_precomp=z(1, 3))
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(means3D=z(1, 3), means2D=z(1, 3), opacities=z(1, 1), colors_precomp=z(1, 3), scales=z(1, 3), rotations=z(1, 4), cov3D_precomp=z(1, 6))


def test_no_cpu_fallback_cpu_tensors_are_rejected():
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    z = torch.zeros
    rs = GaussianRasterizationSettings(16, 16, 1.0, 1.0, z(3), 1.0, torch.eye(4), torch.eye(4), 0, z(3), False, False)
    with pytest.raises(RuntimeError, match="no CPU path"):
        GaussianRasterizer(rs)(means3D=z(1, 3), means2D=z(1, 3), opacities=z(1, 1), colors_precomp=z(1, 3), scales=z(1, 3), rotations=z(1, 4))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "comfyui-3d-pack_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from\s+oracle|import\s+oracle|from\s+\.*gs_oracle|import\s+gs_oracle)", txt, flags=re.M), f


def test_host_camera_matches_reference_fixtures_and_oracle():
    from gs_b200 import camera
    from oracle import gs_oracle as O
    g = np.load(os.path.join(GOLDEN, "ref_camera.npz"))
    for i in range(3):
        W, H, fovy = g[f"dims{i}"]
        fy = np.deg2rad(fovy); fx = 2 * np.arctan(np.tan(fy / 2) * W / H)
        mc = camera.MiniCam(g[f"c2w{i}"], int(W), int(H), fy, fx, 0.01, 100.0, device="cpu")
        assert np.allclose(mc.world_view_transform.numpy(), g[f"wvt{i}"], atol=1e-6)
        assert np.allclose(mc.full_proj_transform.numpy(), g[f"full{i}"], atol=1e-5)
        assert np.allclose(mc.camera_center.numpy(), g[f"center{i}"], atol=1e-7)
    for el, az in [(0, 0), (30, 100), (-40, -170)]:
        assert np.allclose(camera.orbit_camera(el, az, 1.75), O.orbit_camera(el, az, 1.75), atol=1e-6)
    v = camera.orbit_views(8, 1920, 1080)
    st = O.minicam_settings(O.orbit_camera(0, 45.0, 1.75), 1920, 1080, 49.1)
    assert np.allclose(v[1, :16], st.viewmatrix.reshape(-1).numpy(), atol=1e-6)
    assert np.allclose(v[1, 16:32], st.projmatrix.reshape(-1).numpy(), atol=1e-5)
    assert abs(v[1, 38] - st.tanfovx) < 1e-6 and abs(v[1, 39] - st.tanfovy) < 1e-6


def test_gs_ply_wire_format_roundtrip_and_property_order(tmp_path):
    from gs_b200 import ply
    g = torch.Generator().manual_seed(0)
    N, M = 37, 16
    d = dict(xyz=torch.randn(N, 3, generator=g), shs=torch.randn(N, M, 3, generator=g), opacity=torch.randn(N, 1, generator=g),
             scaling=torch.randn(N, 3, generator=g), rotation=torch.randn(N, 4, generator=g))
    path = str(tmp_path / "g.ply")
    ply.write_gs_ply(path, **d)
    head = open(path, "rb").read(4096).split(b"end_header\n")[0].decode().splitlines()
    props = [l.split()[-1] for l in head if l.startswith("property")]
    # mesh_utils.py:333-344 (construct_list_of_gs_attributes) order
    assert props[:9] == ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"]
    assert props[9] == "f_rest_0" and props[9 + 44] == "f_rest_44" and props[-8:] == ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    assert all(l.startswith("property float ") for l in head if l.startswith("property"))
    back = ply.read_gs_ply(path)
    for k in d:
        assert torch.equal(back[k], d[k]), k
    # channel-major SH planes, as to_ply's transpose(1,2).flatten (main_3DGS_renderer.py:475-484)
    rows = ply.pack_rows(**d)
    assert float(rows[5, 9 + 1 * 15 + 3]) == float(d["shs"][5, 4, 1])          # f_rest_{c*(M-1)+k} = coeff k+1 of channel c
    assert ply.max_sh_degree_from_properties(45) == 3 and ply.max_sh_degree_from_properties(0) == 0


def test_switch_axis_and_scale_matches_the_reference_sequence():
    """gs_b200.ply.switch_axis_and_scale against a step-by-step restatement of switch_ply_axis_and_scale
    (mesh_processer/mesh_utils.py:446-472) in numpy, including the in-place column permutation of switch_vector_axis
    (:433-443) and the odd-inversion sign flip; quaternion <-> axis-angle round trip is the identity on unit quaternions."""
    import numpy as np
    import torch
    from gs_b200 import ply
    rng = np.random.RandomState(0)
    n = 257
    q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[0] = [1, 0, 0, 0]; q[1] = [-1, 0, 0, 0]                       # zero vector part: the k = 2 / k = 0.5 branches
    f = dict(xyz=torch.from_numpy(rng.normal(size=(n, 3)).astype(np.float32)), shs=torch.zeros(n, 4, 3), opacity=torch.zeros(n, 1),
             scaling=torch.from_numpy(rng.normal(size=(n, 3)).astype(np.float32)), rotation=torch.from_numpy(q.astype(np.float32)))
    rt = ply.axis_angle_to_quaternion(ply.quaternion_to_axis_angle(f["rotation"]))
    same = torch.minimum((rt - f["rotation"]).abs().amax(1), (rt + f["rotation"]).abs().amax(1))     # q and -q are one rotation
    assert float(same.max()) < 1e-5
    for axis, scale, inv in (([2, 0, 1], [1.0, -1.0, 1.0], 1), ([0, 2, 1], [2.0, 2.0, 2.0], 0), ([0, 1, 2], [1.0, 1.0, 1.0], 2)):
        out = ply.switch_axis_and_scale(f, axis, scale, inv)
        sc = np.asarray(scale, dtype=np.float32)
        xyz = (f["xyz"].numpy() * sc)[:, axis]
        w, v = q[:, 0], q[:, 1:]
        s = np.linalg.norm(v, axis=1)
        two_theta = 2 * np.where(w < 0, np.arctan2(-s, -w), np.arctan2(s, w))
        k = np.where(s > 0, two_theta / np.where(s > 0, s, 1), 2.0)
        aa = ((v * k[:, None]) * sc)[:, axis]
        if inv % 2:
            aa = -aa
        t = np.linalg.norm(aa, axis=1)
        kk = np.where(t > 0, np.sin(0.5 * t) / np.where(t > 0, t, 1), 0.5)
        ref_q = np.concatenate([np.where(t > 0, np.cos(0.5 * t), 1.0)[:, None], aa * kk[:, None]], 1)
        assert np.allclose(out["xyz"].numpy(), xyz, atol=1e-6)
        assert np.array_equal(out["scaling"].numpy(), f["scaling"].numpy()[:, axis])
        assert np.allclose(out["rotation"].numpy(), ref_q, atol=2e-6)
        assert out["shs"] is f["shs"] and out["opacity"] is f["opacity"]


def test_bench_harness_imports_without_a_gpu_and_the_clock_poller_degrades_gracefully():
    """bench.py must be importable on the CPU box (the driver parses its line there), `--impl reference` and the three
    workloads must be selectable, and the nvidia-smi poller must neither raise nor report clocks when the tool is absent
    or prints nothing (clocks = None is what the line then carries, never a made-up figure)."""
    import importlib
    import shutil
    import sys
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    for interval in (20, 100):
        c = bench.Clocks(0, interval)
        out = c.stop()
        assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons"}
        if shutil.which("nvidia-smi") is None:
            assert c.p is None and out["sm_mhz"] is None and out["reasons"] == []
    src = open(bench.__file__).read()
    for token in ("--impl", "--workload", "--gpus", "--steps", "--warmup", '"roofline"', '"cpu_baseline"', '"e2e"', '"gpu_launches"', '"clocks"'):
        assert token in src, token
