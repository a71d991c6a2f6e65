"""Where does each warp role of the conv engine wait?  Needs a profile build:
    AANET_NVCC_DEFS=-DAANET_PROFILE python -m aanet_b200.build --force && python profiles/role_profile.py
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from aanet_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
torch.manual_seed(0)
names = {0: "MMA thread total", 1: "producer: load+compute+store", 2: "producer: wait free stage",
         3: "epilogue: wait accumulator", 4: "epilogue: work", 5: "loader: wait free stage", 6: "loader: issue",
         7: "MMA: wait drained accumulator", 8: "MMA: wait A", 9: "MMA: wait B", 10: "MMA: issue+commit", 11: "tiles"}


def run(tag, fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    for m in range(4):                       # drop what the warm-up calls accumulated
        getattr(lib, "aanet_profile_read_m%d" % m)((ctypes.c_longlong * (148 * 16))())
    fn()
    torch.cuda.synchronize()
    # one counter array per engine translation unit (MODE 0..3); reading clears it, so only the unit whose kernels
    # ran since the last read has tiles > 0
    a = np.zeros((0, 16))
    for m in range(4):
        buf = (ctypes.c_longlong * (148 * 16))()
        assert getattr(lib, "aanet_profile_read_m%d" % m)(buf) == 0
        b = np.ctypeslib.as_array(buf).reshape(148, 16).astype(np.float64)
        a = np.concatenate([a, b[b[:, 11] > 0]])
    print("%s  (%d CTAs, %.1f tiles/CTA)" % (tag, len(a), a[:, 11].mean()))
    for k in (0, 10, 8, 9, 7, 1, 2, 3, 4, 5, 6):
        print("   %-32s %9.0f cycles/CTA   %8.0f per tile" % (names[k], a[:, k].mean(), (a[:, k] / a[:, 11]).mean()))


C, H, W = 64, 128, 416
x = torch.randn(1, H, W, C, device=dev)
wp1 = ops.pack_conv_weight(torch.randn(C, C, 1, 1, device=dev) / 8)
wp3 = ops.pack_conv_weight(torch.randn(C, C, 3, 3, device=dev) / 24)
om = torch.cat([2 * torch.randn(1, H, W, 36, device=dev), torch.rand(1, H, W, 18, device=dev) * 2], -1).contiguous()
run("dense 1x1 64->64", lambda: ops.conv2d_nhwc(x, wp1, C, 1, 1, None, None, None, None, 1))
run("dense 3x3 64->64", lambda: ops.conv2d_nhwc(x, wp3, C, 3, 3, None, None, None, None, 1, 0.2, 1, 1))
run("deform 3x3 64->64", lambda: ops.mdcn_nhwc(x, om, wp3, C, 3, 3, None, None, None, True, 1, 2, 2, 1, 2))
