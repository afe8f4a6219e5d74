from gs_b200.ngp import MLP, trunc_exp  # noqa: F401
