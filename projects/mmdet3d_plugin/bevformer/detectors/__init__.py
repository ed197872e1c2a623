from .bevformer_occ import BEVFormerOcc   # noqa: F401
