"""Import-name shim for the nerfacc surface the reference uses (Instant_NGP.py:30,117,129-149)."""
from gs_b200.ngp import OccGridEstimator, render_weight_from_density, accumulate_along_rays  # noqa: F401
__version__ = "0.5.3+b200"
