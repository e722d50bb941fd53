"""bevipm -- B200-native multi-view IPM warp + BEV fusion (drop-in for the reference's
project/models/fusion/{geometry,fusion}.py).  CUDA only: importing the product path without
libbevipm.so raises."""
from . import rig  # host-side synthetic inputs; no GPU needed
from .modules import (AttentionFusion, ConcatFusion, DeformAttnFusion, FoldedConcatProjIPM, FusedIPM, FusionModule, GeometryTransformer,
                      ImageSpaceDeformAttnFusion, SimpleFusion,
                      pack_calibration)
from . import ops, sharding

__all__ = ["GeometryTransformer", "FusedIPM", "FoldedConcatProjIPM", "FusionModule", "SimpleFusion", "ConcatFusion", "AttentionFusion", "DeformAttnFusion", "ImageSpaceDeformAttnFusion",
           "pack_calibration", "ops", "rig", "sharding"]
