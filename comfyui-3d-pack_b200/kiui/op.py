from gs_b200.ngp import inverse_sigmoid, safe_normalize  # noqa: F401
