"""Drop-in `projects.mmdet3d_plugin` package (config: `plugin_dir = 'projects/mmdet3d_plugin/'`,
bevformer_base_occ.py:6-7; imported by the reference's tools/test.py:137-158).  Importing it registers the
hot-path classes under their reference type names; the implementations live in `occnet_b200.plugin`."""
from .bevformer import *          # noqa: F401,F403
from .datasets import ray_metrics  # noqa: F401
from .datasets.samplers import DistributedSampler  # noqa: F401
