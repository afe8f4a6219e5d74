from gs_b200.ngp import GridEncoder  # noqa: F401
