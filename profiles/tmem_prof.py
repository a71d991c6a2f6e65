"""Role cycle counters (block 0) of the TMEM-A kernels: dense 3x3 / offset head / deformable conv at the 1/3 scale."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from aanet_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
C, H, W = 64, 128, 416
x = torch.randn(1, H, W, C, device=dev)
sc, sh = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
os.environ["AANET_DENSE_TMEM"] = "1"
for name, (w, k, pad, dil, grp, Co) in {
        "3x3": (torch.randn(C, C, 3, 3, device=dev) / 24, 3, 1, 1, 1, C),
        "head": (torch.randn(54, 32, 3, 3, device=dev) / 17, 3, 2, 2, 2, 54)}.items():
    wp = ops.pack_conv_weight(w, grp)
    head = Co == 54
    for i in range(3):
        os.environ["AANET_HALO_PROF"] = "1" if i == 2 else "0"
        ops.conv2d_nhwc(x, wp, Co, k, k, None, None if head else sc, None if head else sh, None,
                        ops.ACT_OFFSET_MASK if head else ops.ACT_RELU, 0.0, 1, pad, dil, grp, out_nchw=head,
                        n_offset_ch=36 if head else 0, mask_scale=2.0)
        torch.cuda.synchronize()
    print("^", name, flush=True)
