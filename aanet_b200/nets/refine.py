"""Refinement modules with the fused front end.

Same constructors, sub-module names and state_dict keys as the reference's nets/refinement.py
(`StereoDRNetRefinement` :60-106, `HourglassRefinement` :109-202) and the building blocks they take from
nets/feature.py (`BasicBlock` :42-76, `BasicConv` :314-339, `Conv2x` :342-376), so reference checkpoints load
with strict=True.  What changes is how forward() starts: upsample + rescale + disp_warp + error + concat
(refinement.py:80-95 / :144-160, warp.py:41-64; ~15 launches and a device synchronisation in the reference) is one
sm_100a kernel (`ops.refine_frontend`) when autograd is off; with autograd on the same arithmetic runs as torch ops
(without the `assert disp.min() >= 0` synchronisation).  The deformable layers of the hourglass are this package's
DeformConv2d, i.e. the tcgen05 engine.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .deform import DeformConv2d


def conv3x3(in_planes, out_planes, stride=1, groups=1, dilation=1):
    """3x3 convolution with padding = dilation (feature.py:15-24, plain variant)."""
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=dilation, groups=groups,
                     bias=False, dilation=dilation)


def conv2d(in_channels, out_channels, kernel_size=3, stride=1, dilation=1, groups=1):
    """conv + BN + LeakyReLU(0.2) (refinement.py:10-16)."""
    return nn.Sequential(nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                                   padding=dilation, dilation=dilation, bias=False, groups=groups),
                         nn.BatchNorm2d(out_channels),
                         nn.LeakyReLU(0.2, inplace=True))


class BasicBlock(nn.Module):
    """feature.py:42-76."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, groups=1, base_width=64, dilation=1,
                 norm_layer=None, leaky_relu=True):
        super(BasicBlock, self).__init__()
        if norm_layer is None:
            norm_layer = nn.BatchNorm2d
        self.conv1 = conv3x3(inplanes, planes, stride=stride, dilation=dilation)
        self.bn1 = norm_layer(planes)
        self.relu = nn.LeakyReLU(0.2, inplace=True) if leaky_relu else nn.ReLU(inplace=True)
        self.conv2 = conv3x3(planes, planes, dilation=dilation)
        self.bn2 = norm_layer(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        identity = x
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        if self.downsample is not None:
            identity = self.downsample(x)
        out = out + identity
        return self.relu(out)


class BasicConv(nn.Module):
    """feature.py:314-339 (2-D variants; the 3-D ones belong to the GANet cost aggregation, out of scope)."""

    def __init__(self, in_channels, out_channels, deconv=False, is_3d=False, bn=True, relu=True, **kwargs):
        super(BasicConv, self).__init__()
        if is_3d:
            raise NotImplementedError("3-D BasicConv is not on the AANet path")
        self.relu = relu
        self.use_bn = bn
        if deconv:
            self.conv = nn.ConvTranspose2d(in_channels, out_channels, bias=False, **kwargs)
        else:
            self.conv = nn.Conv2d(in_channels, out_channels, bias=False, **kwargs)
        self.bn = nn.BatchNorm2d(out_channels)

    def forward(self, x):
        x = self.conv(x)
        if self.use_bn:
            x = self.bn(x)
        if self.relu:
            x = F.relu(x, inplace=True)
        return x


class Conv2x(nn.Module):
    """feature.py:342-376."""

    def __init__(self, in_channels, out_channels, deconv=False, is_3d=False, concat=True, bn=True, relu=True,
                 mdconv=False):
        super(Conv2x, self).__init__()
        if is_3d:
            raise NotImplementedError("3-D Conv2x is not on the AANet path")
        self.concat = concat
        kernel = 4 if deconv else 3
        self.conv1 = BasicConv(in_channels, out_channels, deconv, is_3d, bn=True, relu=True, kernel_size=kernel,
                               stride=2, padding=1)
        if self.concat:
            if mdconv:
                self.conv2 = DeformConv2d(out_channels * 2, out_channels, kernel_size=3, stride=1)
            else:
                self.conv2 = BasicConv(out_channels * 2, out_channels, False, is_3d, bn, relu, kernel_size=3,
                                       stride=1, padding=1)
        else:
            self.conv2 = BasicConv(out_channels, out_channels, False, is_3d, bn, relu, kernel_size=3, stride=1,
                                   padding=1)

    def forward(self, x, rem):
        x = self.conv1(x)
        assert x.size() == rem.size()
        x = torch.cat((x, rem), 1) if self.concat else x + rem
        return self.conv2(x)


def refine_frontend_torch(low_disp, left_img, right_img):
    """The reference's arithmetic as differentiable torch ops (refinement.py:80-95, warp.py:41-64 without the
    synchronising assert and without the unused validity mask)."""
    low = low_disp.unsqueeze(1)
    scale_factor = left_img.size(-1) / low.size(-1)
    if scale_factor == 1.0:
        disp = low
    else:
        disp = F.interpolate(low, size=left_img.size()[-2:], mode='bilinear', align_corners=False) * scale_factor
    b, _, h, w = right_img.size()
    xs = torch.arange(0, w, device=disp.device, dtype=disp.dtype).view(1, 1, 1, w).expand(b, 1, h, w)
    ys = torch.arange(0, h, device=disp.device, dtype=disp.dtype).view(1, 1, h, 1).expand(b, 1, h, w)
    gx = 2 * ((xs - disp) / (w - 1)) - 1
    gy = 2 * (ys / (h - 1)) - 1
    grid = torch.cat((gx, gy), dim=1).permute(0, 2, 3, 1)
    warped = F.grid_sample(right_img, grid, mode='bilinear', padding_mode='border', align_corners=True)
    return torch.cat((warped - left_img, left_img), dim=1), disp


def refine_frontend(low_disp, left_img, right_img):
    """(cat(warped_right - left, left) [B,6,H,W], disp [B,1,H,W]) from low_disp [B,h,w]: one sm_100a launch when
    nothing needs a gradient, the torch composite otherwise."""
    assert low_disp.dim() == 3
    if torch.is_grad_enabled() and (low_disp.requires_grad or left_img.requires_grad or right_img.requires_grad):
        return refine_frontend_torch(low_disp, left_img, right_img)
    return ops.refine_frontend(low_disp, left_img, right_img)


class _RefineBase(nn.Module):
    """Shared stem of both refinement nets: 6 -> 16 channels on cat(error, left), 1 -> 16 on the disparity."""

    def _make_stem(self):
        self.conv1 = conv2d(6, 16)
        self.conv2 = conv2d(1, 16)

    def _stem(self, low_disp, left_img, right_img):
        concat1, disp = refine_frontend(low_disp, left_img, right_img)
        return torch.cat((self.conv1(concat1), self.conv2(disp)), dim=1), disp

    def _finish(self, feat, disp):
        return F.relu(disp + self.final_conv(feat)).squeeze(1)          # [B, H, W]


class StereoDRNetRefinement(_RefineBase):
    """Six dilated residual blocks on the 32-channel stem (reference refinement.py:60-106)."""

    def __init__(self):
        super(StereoDRNetRefinement, self).__init__()
        self._make_stem()
        self.dilation_list = [1, 2, 4, 8, 1, 1]
        self.dilated_blocks = nn.Sequential(*[BasicBlock(32, 32, stride=1, dilation=d) for d in self.dilation_list])
        self.final_conv = nn.Conv2d(32, 1, 3, 1, 1)

    def forward(self, low_disp, left_img, right_img):
        feat, disp = self._stem(low_disp, left_img, right_img)
        return self._finish(self.dilated_blocks(feat), disp)


class HourglassRefinement(_RefineBase):
    """Two stacked hourglasses with deformable convolutions at the 1/8 and 1/16 levels (reference
    refinement.py:109-202; H and W must be divisible by 16).  Layer names follow the reference so that its
    checkpoints load: conv{k}a / deconv{k}a for the first hourglass, conv{k}b / deconv{k}b for the second."""
    widths = (32, 48, 64, 96, 128)

    def __init__(self):
        super(HourglassRefinement, self).__init__()
        self._make_stem()
        w = self.widths
        self.conv_start = DeformConv2d(w[0], w[0])
        for k in (1, 2, 3, 4):
            deform = k >= 3
            if deform:
                down = DeformConv2d(w[k - 1], w[k], kernel_size=3, stride=2)
            else:
                down = BasicConv(w[k - 1], w[k], kernel_size=3, stride=2, padding=1)
            setattr(self, "conv%da" % k, down)
        for k in (4, 3, 2, 1):
            setattr(self, "deconv%da" % k, Conv2x(w[k], w[k - 1], deconv=True))
        for k in (1, 2, 3, 4):
            setattr(self, "conv%db" % k, Conv2x(w[k - 1], w[k], mdconv=k >= 3))
        for k in (4, 3, 2, 1):
            setattr(self, "deconv%db" % k, Conv2x(w[k], w[k - 1], deconv=True))
        self.final_conv = nn.Conv2d(w[0], 1, 3, 1, 1)

    def forward(self, low_disp, left_img, right_img):
        x, disp = self._stem(low_disp, left_img, right_img)
        skips = [self.conv_start(x)]                       # skips[k]: feature at 1 / 2^k of the image size
        for k in (1, 2, 3, 4):
            skips.append(getattr(self, "conv%da" % k)(skips[-1]))
        x = skips[4]
        for k in (4, 3, 2, 1):                             # first decoder refreshes the skips on its way up
            x = skips[k - 1] = getattr(self, "deconv%da" % k)(x, skips[k - 1])
        for k in (1, 2, 3, 4):                             # second encoder consumes and refreshes them again
            x = getattr(self, "conv%db" % k)(x, skips[k])
            if k < 4:
                skips[k] = x
        for k in (4, 3, 2, 1):
            x = getattr(self, "deconv%db" % k)(x, skips[k - 1])
        return self._finish(x, disp)
