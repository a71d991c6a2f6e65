"""Target for `ncu --set full -k regex:corr_umma`: the three correlation launches of the bench workload."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from aanet_b200 import ops  # noqa: E402

torch.manual_seed(0)
for (C, D, H, W) in [(128, 64, 128, 416), (128, 32, 64, 208), (128, 16, 32, 104)]:
    L = torch.relu(torch.randn(1, C, H, W, device="cuda"))
    R = torch.relu(torch.randn(1, C, H, W, device="cuda"))
    for _ in range(2):
        out = ops.correlation_nhwc(L, R, D)
torch.cuda.synchronize()
print("ok", float(out.sum()))
