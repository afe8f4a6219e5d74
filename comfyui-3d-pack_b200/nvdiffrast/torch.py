from gs_b200.meshops import (RasterizeCudaContext, RasterizeGLContext, rasterize, interpolate, texture, antialias,  # noqa: F401
                             antialias_construct_topology_hash)
