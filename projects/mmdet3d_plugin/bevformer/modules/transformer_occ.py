"""`TransformerOcc` under its reference import path (mmdet TRANSFORMER registry)."""
from occnet_b200.plugin.modules import TransformerOcc   # noqa: F401
