import os, sys
sys.path.insert(0, os.getcwd())
import torch
from aanet_b200 import ops
torch.manual_seed(0)
B, C, H, W, dg = 1, 64, 128, 416, 2
x = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
off = (2 * torch.randn(B, dg * 18, H, W, device="cuda")).requires_grad_()
msk = (2 * torch.sigmoid(torch.randn(B, dg * 9, H, W, device="cuda"))).requires_grad_()
w = (torch.randn(C, C, 3, 3, device="cuda") / 24).requires_grad_()
for _ in range(3):
    out = ops.modulated_deform_conv(x, off, msk, w, None, 1, 2, 2, 1, dg)
    out.backward(torch.ones_like(out))
torch.cuda.synchronize()
print("ok")
