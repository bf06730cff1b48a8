"""se3conv3d_b200 -- B200-native (sm_100a) drop-in for the PNEConvLayerRotEquiv hot path of
lisaweijler/SE3Conv3D's point_cloud_lib.

Layout mirrors the reference's operator interface for this path:
  se3conv3d_b200.point_cloud_lib_ops  the five legacy native entry points (custom_ops/ops_list.cpp:19-26)
  se3conv3d_b200.custom_ops           autograd.Function wrappers (custom_ops/__init__.py:1-6)
  se3conv3d_b200.pc                   Pointcloud(RotEquiv), PointHierarchy(RotEquiv), Grid, neighbourhoods
  se3conv3d_b200.layers               IConvLayer, PNEConvLayerRotEquiv (+ factories)
All compute goes through libse3conv3d_b200.so (include/se3conv3d_b200.h); there is no fallback.
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
