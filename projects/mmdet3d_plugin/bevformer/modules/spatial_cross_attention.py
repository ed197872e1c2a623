"""`SpatialCrossAttention`, `MSDeformableAttention3D` under their reference import path (ATTENTION registry)."""
from occnet_b200.plugin.modules import MSDeformableAttention3D, SpatialCrossAttention   # noqa: F401
from occnet_b200 import ops as ext_module                                                # noqa: F401
