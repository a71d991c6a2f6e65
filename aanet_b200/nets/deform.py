"""ISA building blocks -- drop-in for the reference's nets/deform.py.

Same class names, constructor signatures, sub-module names (`deform_conv`, `offset_conv`, `conv1`,
`bn1`, ...) and therefore the same state_dict keys as the reference (deform.py:17-236), so
utils.filter_specific_params still finds `offset_conv.weight/bias` (utils/utils.py:156-169) and
checkpoints load with strict=True.  The dense 1x1/3x3 convolutions and BatchNorm stay torch/cuDNN
modules; the deformable convolution is the sm_100a kernel.
"""
import torch
import torch.nn as nn

from .. import ops
from .deform_conv import DeformConv, ModulatedDeformConv


def conv3x3(in_planes, out_planes, stride=1, groups=1, dilation=1):
    """3x3 convolution, padding = dilation, no bias (deform.py:6-9)."""
    return nn.Conv2d(in_planes, out_planes, 3, stride=stride, padding=dilation, groups=groups,
                     bias=False, dilation=dilation)


def conv1x1(in_planes, out_planes, stride=1):
    """1x1 convolution, no bias (deform.py:12-14)."""
    return nn.Conv2d(in_planes, out_planes, 1, stride=stride, bias=False)


def bn_affine(bn):
    """(scale, shift) of an eval-mode BatchNorm2d, cached until one of its tensors changes."""
    key = (bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version,
           bn.weight.data_ptr(), bn.running_mean.data_ptr())
    hit = getattr(bn, "_aanet_affine", None)
    if hit is not None and hit[0] == key:
        return hit[1], hit[2]
    with torch.no_grad():
        scale = (bn.weight * torch.rsqrt(bn.running_var + bn.eps)).float().contiguous()
        shift = (bn.bias - bn.running_mean * scale).float().contiguous()
    bn._aanet_affine = (key, scale, shift)
    return scale, shift


def _inference_mode(module):
    return (not module.training) and (not torch.is_grad_enabled())


class DeformConv2d(nn.Module):
    """A (modulated) deformable conv layer with its offset/mask predictor (deform.py:17-97).

    `offset_conv` is a grouped (groups = deformable_groups) dilated 3x3 conv with bias, zero-initialised
    so the layer starts as a regular dilated conv (deform.py:70-76).  Its output is split
    POSITIONALLY: the first dg*2*k*k channels are offsets, the rest mask logits (deform.py:82-85) --
    with dg = 2 the offsets of deformable group 1 therefore come half from conv-group 0 and half from
    conv-group 1.  That quirk is part of the trained weights' meaning and is kept as is.
    """

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, dilation=2, groups=1,
                 deformable_groups=2, modulation=True, double_mask=True, bias=False):
        super().__init__()
        self.modulation = modulation
        self.deformable_groups = deformable_groups
        self.kernel_size = kernel_size
        self.double_mask = double_mask

        op = ModulatedDeformConv if modulation else DeformConv
        self.deform_conv = op(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                              padding=dilation, dilation=dilation, groups=groups,
                              deformable_groups=deformable_groups, bias=bias)
        per_point = 3 if modulation else 2
        self.offset_conv = nn.Conv2d(in_channels, deformable_groups * per_point * kernel_size * kernel_size,
                                     kernel_size=kernel_size, stride=stride, padding=dilation,
                                     dilation=dilation, groups=deformable_groups, bias=True)
        nn.init.zeros_(self.offset_conv.weight)
        nn.init.zeros_(self.offset_conv.bias)

    def _offset_and_mask(self, x):
        om = self.offset_conv(x)
        if not self.modulation:
            return om, None
        n_off = self.deformable_groups * 2 * self.kernel_size * self.kernel_size
        mask = om[:, n_off:].sigmoid()
        if self.double_mask:
            mask = mask * 2
        return om[:, :n_off], mask

    def forward(self, x):
        offset, mask = self._offset_and_mask(x)
        if self.modulation:
            return self.deform_conv(x, offset, mask)
        return self.deform_conv(x, offset)

    def forward_fused(self, x, bn, relu=True):
        """Inference only: deformable conv with `bn` (eval statistics) and ReLU folded into the kernel's
        epilogue.  Numerically the same affine map as BatchNorm2d.eval() up to fp32 rounding."""
        offset, mask = self._offset_and_mask(x)
        dc = self.deform_conv
        scale, shift = bn_affine(bn)
        stride = dc.stride[0] if isinstance(dc.stride, tuple) else dc.stride
        pad = dc.padding[0] if isinstance(dc.padding, tuple) else dc.padding
        dil = dc.dilation[0] if isinstance(dc.dilation, tuple) else dc.dilation
        return ops.modulated_deform_conv_fused(x, offset, mask, dc.weight, getattr(dc, "bias", None), stride,
                                               pad, dil, dc.groups, dc.deformable_groups, scale, shift, relu)


class _BottleneckBase(nn.Module):
    """conv1x1-BN-ReLU, <conv2>-BN-ReLU, conv1x1-BN, +identity, ReLU."""

    def _build(self, inplanes, width, out_planes, conv2, stride, downsample, norm_layer):
        if norm_layer is None:
            norm_layer = nn.BatchNorm2d
        self.conv1 = conv1x1(inplanes, width)
        self.bn1 = norm_layer(width)
        self.conv2 = conv2
        self.bn2 = norm_layer(width)
        self.conv3 = conv1x1(width, out_planes)
        self.bn3 = norm_layer(out_planes)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def _mid(self, out):
        return self.relu(self.bn2(self.conv2(out)))

    def forward(self, x):
        identity = x if self.downsample is None else self.downsample(x)
        out = self.relu(self.bn1(self.conv1(x)))
        out = self._mid(out)
        out = self.bn3(self.conv3(out))
        out += identity
        return self.relu(out)


class _DeformMid:
    def _mid(self, out):
        if _inference_mode(self) and isinstance(self.bn2, nn.BatchNorm2d) and self.conv2.modulation:
            return self.conv2.forward_fused(out, self.bn2, relu=True)
        return self.relu(self.bn2(self.conv2(out)))


class DeformBottleneck(_DeformMid, _BottleneckBase):
    """ResNet bottleneck (expansion 4) with a deformable 3x3 (deform.py:100-141; feature extractor)."""
    expansion = 4
    __constants__ = ['downsample']

    def __init__(self, inplanes, planes, stride=1, downsample=None, groups=1, base_width=64, dilation=1,
                 norm_layer=None):
        super().__init__()
        width = int(planes * (base_width / 64.)) * groups
        self._build(inplanes, width, planes * self.expansion, DeformConv2d(width, width, stride=stride),
                    stride, downsample, norm_layer)


class SimpleBottleneck(_BottleneckBase):
    """Bottleneck without channel expansion, plain 3x3 (deform.py:144-184; ISA modules 0-2)."""

    def __init__(self, inplanes, planes, stride=1, downsample=None, groups=1, base_width=64, dilation=1,
                 norm_layer=None):
        super().__init__()
        width = int(planes * (base_width / 64.)) * groups
        self._build(inplanes, width, planes, conv3x3(width, width, stride, groups, dilation), stride,
                    downsample, norm_layer)


class DeformSimpleBottleneck(_DeformMid, _BottleneckBase):
    """The ISA block: SimpleBottleneck with a modulated deformable 3x3 (deform.py:187-236)."""

    def __init__(self, inplanes, planes, stride=1, downsample=None, groups=1, base_width=64,
                 norm_layer=None, mdconv_dilation=2, deformable_groups=2, modulation=True,
                 double_mask=True):
        super().__init__()
        width = int(planes * (base_width / 64.)) * groups
        conv2 = DeformConv2d(width, width, stride=stride, dilation=mdconv_dilation,
                             deformable_groups=deformable_groups, modulation=modulation,
                             double_mask=double_mask)
        self._build(inplanes, width, planes, conv2, stride, downsample, norm_layer)
