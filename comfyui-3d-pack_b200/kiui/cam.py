from gs_b200.camera import orbit_camera  # noqa: F401
