"""Layers on the hot path with the reference's names (point_cloud_lib/point_cloud_lib/layers/__init__.py)."""
from .base import PreProcessModule, IConvLayer, IConvLayerFactory
from .pne_conv_rot_equiv import PNEConvLayerRotEquiv, PNEConvLayerRotEquivFactory
from .pne_conv import PNEConvLayer, PNEConvLayerFactory
from .blocks import NormLayerPC, BatchNormPC, DropPathPC, SkipConnection, Block, ResNetFormer
