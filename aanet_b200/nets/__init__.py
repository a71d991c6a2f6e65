"""Drop-in replacements for the hot-path modules of the reference's `nets` package."""
from .cost import CostVolume, CostVolumePyramid
from .estimation import DisparityEstimation
from .deform import (DeformConv2d, DeformBottleneck, SimpleBottleneck, DeformSimpleBottleneck,
                     conv1x1, conv3x3)
from .deform_conv import (DeformConv, DeformConvPack, ModulatedDeformConv, ModulatedDeformConvPack,
                          deform_conv, modulated_deform_conv)
from .aggregation import AdaptiveAggregationModule, AdaptiveAggregation
from .refine import StereoDRNetRefinement, HourglassRefinement

__all__ = ['CostVolume', 'CostVolumePyramid', 'DisparityEstimation', 'DeformConv2d', 'DeformBottleneck',
           'SimpleBottleneck', 'DeformSimpleBottleneck', 'conv1x1', 'conv3x3', 'DeformConv',
           'DeformConvPack', 'ModulatedDeformConv', 'ModulatedDeformConvPack', 'deform_conv',
           'modulated_deform_conv', 'AdaptiveAggregationModule', 'AdaptiveAggregation',
           'StereoDRNetRefinement', 'HourglassRefinement']
