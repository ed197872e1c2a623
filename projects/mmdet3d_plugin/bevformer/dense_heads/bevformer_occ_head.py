"""`BEVFormerOccHead` under its reference import path (mmdet HEADS registry)."""
from occnet_b200.plugin.modules import BEVFormerOccHead   # noqa: F401
