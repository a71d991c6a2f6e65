"""Deformable convolution modules -- drop-in for nets/deform_conv/deform_conv.py.

Parameter names, shapes and initialisation follow the reference (deform_conv.py:190-239, :304-351)
so its checkpoints load with strict=True; the arithmetic is libaanet_b200.so (aanet_b200.ops).
"""
import math

import torch
import torch.nn as nn
from torch.nn.modules.utils import _pair, _single

from ...ops import (DeformConvFunction, ModulatedDeformConvFunction, deform_conv,  # noqa: F401
                    modulated_deform_conv)


def _uniform_fan_in_(weight, in_channels, kernel_size):
    # U(-1/sqrt(Cin*kh*kw), +...) as deform_conv.py:228-233 / :339-346
    bound = 1.0 / math.sqrt(in_channels * kernel_size[0] * kernel_size[1])
    with torch.no_grad():
        weight.uniform_(-bound, bound)


class _DeformBase(nn.Module):
    def _common(self, in_channels, out_channels, kernel_size, groups, deformable_groups):
        assert in_channels % groups == 0, \
            'in_channels {} cannot be divisible by groups {}'.format(in_channels, groups)
        assert out_channels % groups == 0, \
            'out_channels {} cannot be divisible by groups {}'.format(out_channels, groups)
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _pair(kernel_size)
        self.groups = groups
        self.deformable_groups = deformable_groups
        self.transposed = False              # nn.Conv2d compatibility, as in the reference
        self.output_padding = _single(0)
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels // groups, *self.kernel_size))


class DeformConv(_DeformBase):
    """DCNv1: forward(x, offset).  stride/padding/dilation are stored as pairs (deform_conv.py:215-217)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deformable_groups=1, bias=False):
        super().__init__()
        assert not bias
        self._common(in_channels, out_channels, kernel_size, groups, deformable_groups)
        self.stride, self.padding, self.dilation = _pair(stride), _pair(padding), _pair(dilation)
        self.reset_parameters()

    def reset_parameters(self):
        _uniform_fan_in_(self.weight, self.in_channels, self.kernel_size)

    def forward(self, x, offset):
        return deform_conv(x, offset, self.weight, self.stride, self.padding, self.dilation, self.groups,
                           self.deformable_groups)


class ModulatedDeformConv(_DeformBase):
    """DCNv2: forward(x, offset, mask).  stride/padding/dilation stay scalars (deform_conv.py:320-322)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deformable_groups=1, bias=True):
        super().__init__()
        self._common(in_channels, out_channels, kernel_size, groups, deformable_groups)
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.with_bias = bias
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        _uniform_fan_in_(self.weight, self.in_channels, self.kernel_size)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, x, offset, mask):
        return modulated_deform_conv(x, offset, mask, self.weight, self.bias, self.stride, self.padding,
                                     self.dilation, self.groups, self.deformable_groups)


def _rename_legacy_offset_keys(state_dict, prefix, local_metadata):
    # deform_conv.py:283-296 / :400-413: pre-v2 checkpoints call the layer `<name>_offset`
    if local_metadata.get('version', None) is None or local_metadata.get('version') < 2:
        for leaf in ('weight', 'bias'):
            new, old = prefix + 'conv_offset.' + leaf, prefix[:-1] + '_offset.' + leaf
            if new not in state_dict and old in state_dict:
                state_dict[new] = state_dict.pop(old)


class DeformConvPack(DeformConv):
    """DeformConv that predicts its own offsets with a zero-initialised conv (deform_conv.py:242-301)."""
    _version = 2

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        k = self.kernel_size
        self.conv_offset = nn.Conv2d(self.in_channels, self.deformable_groups * 2 * k[0] * k[1],
                                     kernel_size=k, stride=_pair(self.stride), padding=_pair(self.padding),
                                     bias=True)
        self.init_offset()

    def init_offset(self):
        nn.init.zeros_(self.conv_offset.weight)
        nn.init.zeros_(self.conv_offset.bias)

    def forward(self, x):
        return deform_conv(x, self.conv_offset(x), self.weight, self.stride, self.padding, self.dilation,
                           self.groups, self.deformable_groups)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, *rest):
        _rename_legacy_offset_keys(state_dict, prefix, local_metadata)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, *rest)


class ModulatedDeformConvPack(ModulatedDeformConv):
    """ModulatedDeformConv with its own offset/mask conv; mask = sigmoid (deform_conv.py:354-418)."""
    _version = 2

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        k = self.kernel_size
        self.conv_offset = nn.Conv2d(self.in_channels, self.deformable_groups * 3 * k[0] * k[1],
                                     kernel_size=k, stride=_pair(self.stride), padding=_pair(self.padding),
                                     bias=True)
        self.init_offset()

    def init_offset(self):
        nn.init.zeros_(self.conv_offset.weight)
        nn.init.zeros_(self.conv_offset.bias)

    def forward(self, x):
        o1, o2, mask = torch.chunk(self.conv_offset(x), 3, dim=1)
        offset = torch.cat((o1, o2), dim=1)
        return modulated_deform_conv(x, offset, torch.sigmoid(mask), self.weight, self.bias, self.stride,
                                     self.padding, self.dilation, self.groups, self.deformable_groups)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, *rest):
        _rename_legacy_offset_keys(state_dict, prefix, local_metadata)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, *rest)
