"""A/B of the TMA-staged correlation kernel against the round-1 register-staged one at the three pyramid scales of
config 2 (C = 128), inputs rotated over sets larger than L2."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from aanet_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
for (H, W, D) in ((128, 416, 64), (64, 208, 32), (32, 104, 16)):
    n = 6
    Ls = [torch.relu(torch.randn(1, 128, H, W, device=dev)) for _ in range(n)]
    for nhwc in (False, True):
        fn = (lambda i: ops.correlation_nhwc(Ls[i], Ls[(i + 1) % n], D)) if nhwc else \
             (lambda i: ops.correlation(Ls[i], Ls[(i + 1) % n], D))
        res = {}
        for flag in ("0", "1"):
            os.environ["AANET_CORR_TMA"] = flag
            res[flag] = bench._timed(fn, n, 24, dev) * 1e3
        alg = 4.0 * H * W * (2 * 128 + D)
        print("%3dx%3d D=%2d %s: register-staged %5.1f us, TMA %5.1f us (%.2f of the %.1f us HBM floor)"
              % (H, W, D, "NHWC" if nhwc else "NCHW", res["0"], res["1"], alg / 6444.4e3 / res["1"], alg / 6444.4e3))
