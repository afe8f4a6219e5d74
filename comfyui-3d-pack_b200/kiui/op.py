from gs_b200.ngp import inverse_sigmoid, safe_normalize  # noqa: F401
from gs_b200.texops import uv_padding  # noqa: F401


def dot(x, y):
    """kiui.op.dot: sum(x * y) over the last axis, kept (mesh_processer/mesh.py:487)."""
    import torch
    return torch.sum(x * y, -1, keepdim=True)
