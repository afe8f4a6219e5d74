"""Import-name shim for the parts of `kiui` the reference's renderer paths import:
kiui.gridencoder.GridEncoder, kiui.nn.{MLP,trunc_exp}, kiui.op.{inverse_sigmoid,safe_normalize}, kiui.cam.orbit_camera."""
__version__ = "0.2.14+b200"
