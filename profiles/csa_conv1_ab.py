"""csa_conv1_nhwc (resize-and-sum + LeakyReLU inside conv1's launch) against csa_fuse_nhwc + the 1x1 convolution, at the
three scales' shapes of config 2."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from aanet_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(3)
n = 6
for name, C, sizes in (("row 0 (1/3 scale)", 64, [(128, 416), (64, 208), (32, 104)]),
                       ("row 1 (1/6 scale)", 32, [(64, 208), (64, 208), (32, 104)])):
    sets = [[torch.randn(1, h, w, C, device=dev) for h, w in sizes] for _ in range(n)]
    wp = ops.pack_conv_weight(torch.randn(C, C, 1, 1, device=dev) / C ** 0.5)
    sc, sh = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
    t_f = bench._timed(lambda i: ops.csa_conv1_nhwc(sets[i], 0.2, wp, C, sc, sh, ops.ACT_RELU), n, 24, dev) * 1e3
    t_a = bench._timed(lambda i: ops.csa_fuse_nhwc(sets[i], 0.2), n, 24, dev) * 1e3
    xs = [ops.csa_fuse_nhwc(s, 0.2) for s in sets]
    t_b = bench._timed(lambda i: ops.conv2d_nhwc(xs[i], wp, C, 1, 1, None, sc, sh, None, ops.ACT_RELU, 0.0, 1, 0, 1, 1), n, 24, dev) * 1e3
    print("%-18s fused %.1f us | csa_fuse %.1f us + conv1 %.1f us" % (name, t_f, t_a, t_b), flush=True)
