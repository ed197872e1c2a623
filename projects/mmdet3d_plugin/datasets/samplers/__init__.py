from .distributed_sampler import DistributedSampler   # noqa: F401
