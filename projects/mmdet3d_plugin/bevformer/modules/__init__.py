from .spatial_cross_attention import SpatialCrossAttention, MSDeformableAttention3D   # noqa: F401
from .temporal_self_attention import TemporalSelfAttention                             # noqa: F401
from .encoder import BEVFormerEncoder, BEVFormerLayer                                  # noqa: F401
from .transformer_occ import TransformerOcc                                            # noqa: F401
