"""kiui.mesh_utils names used around the mesh path (MVs_Algorithms/DiffRastMesh/diff_mesh.py:7-8)."""
from gs_b200.meshreg import laplacian_smooth_loss, normal_consistency  # noqa: F401


def _needs_pymeshlab(name):
    def fn(*a, **k):
        raise NotImplementedError(f"kiui.mesh_utils.{name} wraps pymeshlab re-meshing; it is outside the rendering hot path "
                                  f"(SURVEY.md section 8) and pymeshlab is not available in this environment")
    fn.__name__ = name
    return fn


clean_mesh = _needs_pymeshlab("clean_mesh")
decimate_mesh = _needs_pymeshlab("decimate_mesh")
