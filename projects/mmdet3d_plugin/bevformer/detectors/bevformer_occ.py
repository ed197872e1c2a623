"""`BEVFormerOcc` under its reference import path (mmdet DETECTORS registry)."""
from occnet_b200.plugin.modules import BEVFormerOcc   # noqa: F401
