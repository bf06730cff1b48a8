"""autograd.Function wrappers with the reference's names and argument meaning
(point_cloud_lib/point_cloud_lib/custom_ops/__init__.py:1-6), backed by the C ABI."""
from .functions import (FeatBasisProj, BallQuery, KNNQuery, ComputeKeys, RotEquivConv, GammaSkip, FramePool, BatchPool,
                        POOL_MODES)

__all__ = ["FeatBasisProj", "BallQuery", "KNNQuery", "ComputeKeys", "RotEquivConv", "GammaSkip", "FramePool", "BatchPool",
           "POOL_MODES"]
