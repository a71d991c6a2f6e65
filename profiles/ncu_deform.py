"""Three launches of the 1/3-scale deformable conv through the channels-last engine call (for ncu).
    python profiles/ncu_deform.py [offset sigma px]      (AANET_DEFORM_HALO=1 selects the halo-staged kernel)"""
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
sc, sh = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
sigma = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
om = torch.cat([sigma * torch.randn(1, 36, H, W, device=dev), torch.rand(1, 18, H, W, device=dev) * 2], 1).contiguous()
for _ in range(3):
    ops.mdcn_nhwc(x, om, wp3, C, 3, 3, None, sc, sh, True, 1, 2, 2, 1, 2, om_nchw=True)
torch.cuda.synchronize()
print("ok")
