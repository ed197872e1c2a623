"""`TemporalSelfAttention` under its reference import path (ATTENTION registry)."""
from occnet_b200.plugin.modules import TemporalSelfAttention   # noqa: F401
from occnet_b200 import ops as ext_module                       # noqa: F401
