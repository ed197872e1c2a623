from .bevformer_occ_head import BEVFormerOccHead   # noqa: F401
