"""Three launches of the 1/3-scale deformable conv through the channels-last engine call (for ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from aanet_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
C, H, W = 64, 128, 416
x = torch.randn(1, H, W, C, device=dev)
wp3 = ops.pack_conv_weight(torch.randn(C, C, 3, 3, device=dev) / 24)
sigma = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
om = torch.cat([sigma * torch.randn(1, H, W, 36, device=dev), torch.rand(1, H, W, 18, device=dev) * 2], -1).contiguous()
for _ in range(3):
    ops.mdcn_nhwc(x, om, wp3, C, 3, 3, None, None, None, True, 1, 2, 2, 1, 2)
    ops.conv2d_nhwc(x, wp3, C, 3, 3, None, None, None, None, 1, 0.2, 1, 1)
torch.cuda.synchronize()
print("ok")
