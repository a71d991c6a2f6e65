"""Correlation cost volume modules -- drop-in for the reference's nets/cost.py.

`CostVolume(max_disp, feature_similarity='correlation')` and `CostVolumePyramid(...)` keep the
reference constructors and call conventions (nets/cost.py:5-76).  The correlation branch (cost.py:40-48) is
the hot path (tcgen05 banded GEMM); 'difference' / 'concat' (cost.py:22-38), the 5-D volumes of the
StereoNet/PSMNet/GC-Net style variants, are one memory-bound kernel each.
"""
import torch
import torch.nn as nn

from .. import ops
from ..streams import fork_join


class CostVolume(nn.Module):
    def __init__(self, max_disp, feature_similarity='correlation'):
        super().__init__()
        self.max_disp = max_disp
        self.feature_similarity = feature_similarity

    def forward(self, left_feature, right_feature):
        if self.feature_similarity in ('difference', 'concat'):
            return ops.cost_volume_5d(left_feature, right_feature, self.max_disp, self.feature_similarity)
        if self.feature_similarity != 'correlation':
            raise NotImplementedError          # cost.py:50-51
        return ops.correlation(left_feature, right_feature, self.max_disp)


class CostVolumePyramid(nn.Module):
    """One correlation volume per pyramid level, level s using max_disp // 2**s (cost.py:64-76)."""

    def __init__(self, max_disp, feature_similarity='correlation'):
        super().__init__()
        self.max_disp = max_disp
        self.feature_similarity = feature_similarity

    def forward(self, left_feature_pyramid, right_feature_pyramid):
        pairs = list(zip(left_feature_pyramid, right_feature_pyramid))
        if self.feature_similarity != 'correlation':      # cost.py:64-76 builds a CostVolume per level
            return [CostVolume(self.max_disp // (2 ** s), self.feature_similarity)(l, r)
                    for s, (l, r) in enumerate(pairs)]
        if pairs and pairs[0][0].is_cuda and not torch.is_grad_enabled():
            # inference: the scales are independent, run them on parallel streams
            return fork_join(pairs[0][0].device,
                             [(lambda s=s, l=l, r=r: ops.correlation(l, r, self.max_disp // (2 ** s)))
                              for s, (l, r) in enumerate(pairs)])
        return [ops.correlation(l, r, self.max_disp // (2 ** s)) for s, (l, r) in enumerate(pairs)]
