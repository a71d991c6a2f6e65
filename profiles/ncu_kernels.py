"""A few launches of each hot kernel for `ncu --set full` (run under gpurun)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from aanet_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
C, H, W = 64, 128, 416
x = torch.randn(1, C, H, W, device=dev)
w = torch.randn(C, C, 3, 3, device=dev) / 24
m = 2 * torch.sigmoid(torch.randn(1, 18, H, W, device=dev))
for sigma in (2.0, 0.3):
    off = sigma * torch.randn(1, 36, H, W, device=dev)
    for _ in range(2):
        ops.modulated_deform_conv(x, off, m, w, None, 1, 2, 2, 1, 2)
w1 = torch.randn(C, C, 1, 1, device=dev) / 8
for _ in range(2):
    ops.conv2d_fused(x, w1, None, None, None, None, 1, 0.0, 1, 0, 1, 1)
    ops.conv2d_fused(x, w, None, None, None, None, 1, 0.0, 1, 1, 1, 1)
L = torch.relu(torch.randn(1, 128, H, W, device=dev))
R = torch.relu(torch.randn(1, 128, H, W, device=dev))
for _ in range(2):
    ops.correlation(L, R, 64)
    ops.soft_argmin(x, True)
torch.cuda.synchronize()
print("ok")
