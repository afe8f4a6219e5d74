"""kiui.typing: star-import of the typing names plus the two array types (mesh_processer/mesh.py:9)."""
from typing import *  # noqa: F401,F403

from numpy import ndarray  # noqa: F401
from torch import Tensor  # noqa: F401
