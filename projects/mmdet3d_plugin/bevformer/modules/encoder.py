"""`BEVFormerEncoder`, `BEVFormerLayer` under their reference import path (registered in
TRANSFORMER_LAYER_SEQUENCE / TRANSFORMER_LAYER); implementation: occnet_b200/plugin/modules.py."""
from occnet_b200.plugin.modules import BEVFormerEncoder, BEVFormerLayer   # noqa: F401
from occnet_b200 import ops as ext_module                                  # the reference binds mmcv `_ext` here
