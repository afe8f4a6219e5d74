"""Import-name shim for `simple_knn` (main_3DGS_renderer.py:408: `from simple_knn._C import distCUDA2`)."""
