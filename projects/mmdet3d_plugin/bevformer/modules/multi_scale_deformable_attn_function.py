"""Operator wrappers under their reference import path: same names, argument order and error behaviour as the
autograd Functions that bind `mmcv._ext.ms_deform_attn_forward / _backward`; arithmetic in libocc_b200 (C ABI)."""
from occnet_b200.ops import (MultiScaleDeformableAttnFunction_fp16, MultiScaleDeformableAttnFunction_fp32,   # noqa: F401
                             ms_deform_attn_backward, ms_deform_attn_forward)
