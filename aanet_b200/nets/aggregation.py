"""Adaptive aggregation (ISA + CSA) -- drop-in for the AdaptiveAggregation* classes of the reference's
nets/aggregation.py (:313-464).

Module tree and state_dict keys are the reference's (`fusions.<i>.branches.<s>.<blk>`,
`fusions.<i>.fuse_layers.<i>.<j>...`, `final_conv.<s>`).  The exchange convolutions stay torch/cuDNN
modules; the resize + ordered sum + LeakyReLU tail of each output scale (aggregation.py:387-400) is one
sm_100a kernel (aanet_b200.ops.csa_fuse).  The 3-D-conv aggregators that share the reference file
(StereoNet/PSMNet/GC-Net, :70-309) are not part of AANet and are out of scope.
"""
import torch.nn as nn

import torch

from .. import ops
from .deform import SimpleBottleneck, DeformSimpleBottleneck, _inference_mode


def _conv_bn(cin, cout, k, stride=1, pad=0, act=False):
    layers = [nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=pad, bias=False),
              nn.BatchNorm2d(cout)]
    if act:
        layers.append(nn.LeakyReLU(0.2, inplace=True))
    return nn.Sequential(*layers)


class AdaptiveAggregationModule(nn.Module):
    def __init__(self, num_scales, num_output_branches, max_disp, num_blocks=1, simple_bottleneck=False,
                 deformable_groups=2, mdconv_dilation=2):
        super().__init__()
        self.num_scales = num_scales
        self.num_output_branches = num_output_branches
        self.max_disp = max_disp
        self.num_blocks = num_blocks

        def width(s):
            return max_disp // (2 ** s)

        # intra-scale aggregation: one stack of bottlenecks per scale (aggregation.py:331-344)
        self.branches = nn.ModuleList()
        for s in range(num_scales):
            blocks = []
            for _ in range(num_blocks):
                if simple_bottleneck:
                    blocks.append(SimpleBottleneck(width(s), width(s)))
                else:
                    blocks.append(DeformSimpleBottleneck(width(s), width(s), modulation=True,
                                                         mdconv_dilation=mdconv_dilation,
                                                         deformable_groups=deformable_groups))
            self.branches.append(nn.Sequential(*blocks))

        # cross-scale aggregation: fuse_layers[i][j] maps scale j to scale i (aggregation.py:346-371)
        self.fuse_layers = nn.ModuleList()
        for i in range(num_output_branches):
            row = nn.ModuleList()
            for j in range(num_scales):
                if i == j:
                    row.append(nn.Identity())
                elif i < j:     # coarser -> finer: 1x1 conv + BN at the coarse size, resized later
                    row.append(_conv_bn(width(j), width(i), 1))
                else:           # finer -> coarser: (i-j) stride-2 3x3 convs, LeakyReLU between them
                    chain = [_conv_bn(width(j), width(j), 3, 2, 1, act=True) for _ in range(i - j - 1)]
                    chain.append(_conv_bn(width(j), width(i), 3, 2, 1))
                    row.append(nn.Sequential(*chain))
            self.fuse_layers.append(row)

        self.relu = nn.LeakyReLU(0.2, inplace=True)

    def forward(self, x):
        assert len(self.branches) == len(x)
        for s, branch in enumerate(self.branches):
            for blk in range(self.num_blocks):
                x[s] = branch[blk](x[s])       # in place on the caller's list, like aggregation.py:378-382
        if self.num_scales == 1:
            return x
        slope = self.relu.negative_slope
        return [ops.csa_fuse([self.fuse_layers[i][j](x[j]) for j in range(len(self.branches))], slope)
                for i in range(len(self.fuse_layers))]


class AdaptiveAggregation(nn.Module):
    def __init__(self, max_disp, num_scales=3, num_fusions=6, num_stage_blocks=1, num_deform_blocks=2,
                 intermediate_supervision=True, deformable_groups=2, mdconv_dilation=2):
        super().__init__()
        self.max_disp = max_disp
        self.num_scales = num_scales
        self.num_fusions = num_fusions
        self.intermediate_supervision = intermediate_supervision

        stages = []
        for i in range(num_fusions):
            last = i == num_fusions - 1
            n_out = num_scales if (intermediate_supervision or not last) else 1
            stages.append(AdaptiveAggregationModule(num_scales=num_scales, num_output_branches=n_out,
                                                    max_disp=max_disp, num_blocks=num_stage_blocks,
                                                    mdconv_dilation=mdconv_dilation,
                                                    deformable_groups=deformable_groups,
                                                    simple_bottleneck=i < num_fusions - num_deform_blocks))
        self.fusions = nn.Sequential(*stages)

        # 1x1 conv WITH bias per kept scale (aggregation.py:443-450)
        self.final_conv = nn.ModuleList()
        for s in range(num_scales):
            c = max_disp // (2 ** s)
            self.final_conv.append(nn.Conv2d(c, c, kernel_size=1))
            if not intermediate_supervision:
                break

    use_fused_inference = True     # eval + no_grad: channels-last tcgen05 executor (aanet_b200/fused.py)

    def forward(self, cost_volume):
        assert isinstance(cost_volume, list)
        if self.use_fused_inference and _inference_mode(self) and cost_volume[0].is_cuda \
                and cost_volume[0].dtype == torch.float32:
            from .. import fused
            if fused.supported(self):
                return fused.run(self, cost_volume)
        for i in range(self.num_fusions):
            cost_volume = self.fusions[i](cost_volume)
        return [conv(cost_volume[s]) for s, conv in enumerate(self.final_conv)]
