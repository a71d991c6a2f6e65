"""Per-kernel timings at the bench workload's shapes (CUDA events, inputs rotated beyond L2).

    python profiles/microbench.py            # prints one line per kernel, JSON at the end
Compares: aanet_b200 tcgen05 mdconv vs its FFMA kernel vs the reference's own CUDA op (oracle/_ref, if
built) ; conv2d_fused vs cuDNN fp32 (+BN+ReLU) ; correlation / soft-argmin / CSA fuse vs their rooflines.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from aanet_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
HBM = 6444.4
try:
    HBM = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, n_sets, iters=24, warm=3):
    """Average device time per call: the calls are captured into one CUDA graph (no Python / launch
    overhead between kernels) that cycles through n_sets input sets, and the graph is replayed."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side), torch.no_grad():
        for i in range(max(warm, n_sets)):
            fn(i % n_sets)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g), torch.no_grad():
        for i in range(iters):
            fn(i % n_sets)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


results = {}


def report(name, us, mbytes=None, gflop=None):
    s = "%-44s %9.1f us" % (name, us)
    if mbytes is not None:
        s += "  %7.1f GB/s (%4.1f%% HBM)" % (mbytes / us * 1e3, 100 * mbytes / us * 1e3 / HBM)
    if gflop is not None:
        s += "  %7.2f TFLOP/s" % (gflop / us * 1e3)
    print(s, flush=True)
    results[name] = {"us": us, "mb": mbytes, "gflop": gflop}


def rnd(*shape):
    return torch.randn(*shape, device=dev)


# ------------------------------------------------------------------ mdconv at the three scales
ref_op = None
try:
    from oracle import build_ref
    ref_op = build_ref.load()
except Exception as e:
    print("reference op not loadable:", e)

for (C, H, W) in [(64, 128, 416), (32, 64, 208), (16, 32, 104)]:
    n = 6 if C == 64 else 12
    xs = [rnd(1, C, H, W) for _ in range(n)]
    offs = [2 * rnd(1, 36, H, W) for _ in range(n)]
    ms = [2 * torch.sigmoid(rnd(1, 18, H, W)) for _ in range(n)]
    w = rnd(C, C, 3, 3) / (C * 9) ** 0.5
    gf = 2.0 * C * C * 9 * H * W / 1e9
    mb = 4.0 * (2 * C * H * W + 54 * H * W) / 1e6
    ops.FORCE_GENERIC_MDCN = False
    report("mdconv tcgen05   C=%d %dx%d" % (C, H, W),
           timeit(lambda i: ops.modulated_deform_conv(xs[i], offs[i], ms[i], w, None, 1, 2, 2, 1, 2), n), mb, gf)
    soffs = [0.3 * rnd(1, 36, H, W) for _ in range(n)]
    report("mdconv tcgen05 (offsets 0.3*randn) C=%d" % C,
           timeit(lambda i: ops.modulated_deform_conv(xs[i], soffs[i], ms[i], w, None, 1, 2, 2, 1, 2), n), mb, gf)
    ops.FORCE_GENERIC_MDCN = True
    report("mdconv FFMA      C=%d %dx%d" % (C, H, W),
           timeit(lambda i: ops.modulated_deform_conv(xs[i], offs[i], ms[i], w, None, 1, 2, 2, 1, 2), n), mb, gf)
    ops.FORCE_GENERIC_MDCN = False
    if ref_op is not None:
        out = torch.empty(1, C, H, W, device=dev)
        bufs = [torch.empty(0, device=dev), torch.empty(0, device=dev)]
        fake_b = torch.empty(1, device=dev)

        def ref_call(i):
            ref_op.modulated_deform_conv_cuda_forward(xs[i], w, fake_b, bufs[0], offs[i], ms[i], out, bufs[1],
                                                      3, 3, 1, 1, 2, 2, 2, 2, 1, 2, False)
        report("mdconv REFERENCE C=%d %dx%d" % (C, H, W), timeit(ref_call, n), mb, gf)
        mine = ops.modulated_deform_conv(xs[0], offs[0], ms[0], w, None, 1, 2, 2, 1, 2)
        ref_call(0)
        print("    max rel diff vs reference op: %.2e" % ((mine - out).abs().max() / out.abs().max()).item())

# ------------------------------------------------------------------ dense convs vs cuDNN
for name, (Ci, Co, H, W, k, st, pad, dil, grp) in {
        "1x1 64->64 128x416": (64, 64, 128, 416, 1, 1, 0, 1, 1),
        "3x3 64->64 128x416": (64, 64, 128, 416, 3, 1, 1, 1, 1),
        "offset 3x3d2g2 64->54 128x416": (64, 54, 128, 416, 3, 1, 2, 2, 2),
        "3x3s2 64->32 128x416": (64, 32, 128, 416, 3, 2, 1, 1, 1),
        "1x1 32->64 64x208": (32, 64, 64, 208, 1, 1, 0, 1, 1),
        "1x1 32->32 64x208": (32, 32, 64, 208, 1, 1, 0, 1, 1),
        "3x3 32->32 64x208": (32, 32, 64, 208, 3, 1, 1, 1, 1),
        "3x3 16->16 32x104": (16, 16, 32, 104, 3, 1, 1, 1, 1)}.items():
    n = 8
    xs = [rnd(1, Ci, H, W) for _ in range(n)]
    w = rnd(Co, Ci // grp, k, k) / (Ci * k * k) ** 0.5
    sc, sh = torch.rand(Co, device=dev) + 0.5, rnd(Co)
    Ho, Wo = (H + 2 * pad - dil * (k - 1) - 1) // st + 1, (W + 2 * pad - dil * (k - 1) - 1) // st + 1
    gf = 2.0 * Co * (Ci // grp) * k * k * Ho * Wo / 1e9
    mb = 4.0 * (Ci * H * W + Co * Ho * Wo) / 1e6
    report("conv tcgen05 +affine+relu " + name,
           timeit(lambda i: ops.conv2d_fused(xs[i], w, None, sc, sh, None, ops.ACT_RELU, 0.0, st, pad, dil, grp), n),
           mb, gf)
    bn = torch.nn.BatchNorm2d(Co).to(dev).eval()
    report("cuDNN conv only           " + name, timeit(lambda i: F.conv2d(xs[i], w, None, st, pad, dil, grp), n),
           mb, gf)
    with torch.no_grad():
        report("cuDNN conv+BN+ReLU        " + name,
               timeit(lambda i: torch.relu_(bn(F.conv2d(xs[i], w, None, st, pad, dil, grp))), n), mb, gf)

# ------------------------------------------------------------------ channels-last engine calls (packed weights)
for name, (Ci, Co, H, W, k, st, pad, dil, grp) in {
        "1x1 64->64 128x416": (64, 64, 128, 416, 1, 1, 0, 1, 1),
        "3x3 64->64 128x416": (64, 64, 128, 416, 3, 1, 1, 1, 1),
        "offset 3x3d2g2 64->54 128x416": (64, 54, 128, 416, 3, 1, 2, 2, 2),
        "3x3s2 64->32 128x416": (64, 32, 128, 416, 3, 2, 1, 1, 1),
        "1x1 32->32 64x208": (32, 32, 64, 208, 1, 1, 0, 1, 1),
        "3x3 32->32 64x208": (32, 32, 64, 208, 3, 1, 1, 1, 1),
        "3x3 16->16 32x104": (16, 16, 32, 104, 3, 1, 1, 1, 1)}.items():
    n = 8
    xs = [rnd(1, H, W, Ci) for _ in range(n)]
    wp = ops.pack_conv_weight(rnd(Co, Ci // grp, k, k) / (Ci * k * k) ** 0.5, grp)
    sc, sh = torch.rand(Co, device=dev) + 0.5, rnd(Co)
    Ho, Wo = (H + 2 * pad - dil * (k - 1) - 1) // st + 1, (W + 2 * pad - dil * (k - 1) - 1) // st + 1
    report("NHWC conv +affine+relu    " + name,
           timeit(lambda i: ops.conv2d_nhwc(xs[i], wp, Co, k, k, None, sc, sh, None, ops.ACT_RELU, 0.0, st, pad, dil, grp), n),
           4.0 * (Ci * H * W + Co * Ho * Wo) / 1e6, 2.0 * Co * (Ci // grp) * k * k * Ho * Wo / 1e9)
for (C, H, W) in [(64, 128, 416), (32, 64, 208), (16, 32, 104)]:
    n = 6 if C == 64 else 12
    xs = [rnd(1, H, W, C) for _ in range(n)]
    oms = [torch.cat([2 * rnd(1, H, W, 36), 2 * torch.sigmoid(rnd(1, H, W, 18))], -1).contiguous() for _ in range(n)]
    wp = ops.pack_conv_weight(rnd(C, C, 3, 3) / (C * 9) ** 0.5)
    report("NHWC mdconv (offsets 2*randn) C=%d" % C,
           timeit(lambda i: ops.mdcn_nhwc(xs[i], oms[i], wp, C, 3, 3, None, None, None, True, 1, 2, 2, 1, 2), n),
           4.0 * (2 * C * H * W + 54 * H * W) / 1e6, 2.0 * C * C * 9 * H * W / 1e9)

# ------------------------------------------------------------------ memory-bound kernels
for s, (C, D, H, W) in enumerate([(128, 64, 128, 416), (128, 32, 64, 208), (128, 16, 32, 104)]):
    n = 4 if s == 0 else 16
    Ls = [torch.relu(rnd(1, C, H, W)) for _ in range(n)]
    Rs = [torch.relu(rnd(1, C, H, W)) for _ in range(n)]
    report("correlation s%d" % s, timeit(lambda i: ops.correlation(Ls[i], Rs[i], D), n),
           4.0 * H * W * (2 * C + D) / 1e6, 2.0 * C * H * (W * D - D * (D - 1) / 2) / 1e9)
n = 12
cs = [rnd(1, 64, 128, 416) for _ in range(n)]
report("soft-argmin s0", timeit(lambda i: ops.soft_argmin(cs[i], True), n), 4.0 * 128 * 416 * 65 / 1e6)
t0 = [rnd(1, 64, 128, 416) for _ in range(n)]
t1 = [rnd(1, 64, 64, 208) for _ in range(n)]
t2 = [rnd(1, 64, 32, 104) for _ in range(n)]
report("CSA fuse out0 (x0 + up2 + up4)", timeit(lambda i: ops.csa_fuse([t0[i], t1[i], t2[i]], 0.2), n),
       4.0 * 64 * (2 * 128 * 416 + 64 * 208 + 32 * 104) / 1e6)
h0 = [rnd(1, 128, 416, 64) for _ in range(n)]
h1 = [rnd(1, 64, 208, 64) for _ in range(n)]
h2 = [rnd(1, 32, 104, 64) for _ in range(n)]
report("CSA fuse out0, channels-last (tiled)", timeit(lambda i: ops.csa_fuse_nhwc([h0[i], h1[i], h2[i]], 0.2), n),
       4.0 * 64 * (2 * 128 * 416 + 64 * 208 + 32 * 104) / 1e6)
# refinement front end at KITTI size: fused kernel vs the torch composite (reference arithmetic, no sync)
from aanet_b200.nets.refine import refine_frontend_torch  # noqa: E402
lows = [torch.rand(1, 128, 416, device=dev) * 60 for _ in range(n)]
imgs = [(torch.rand(1, 3, 384, 1248, device=dev), torch.rand(1, 3, 384, 1248, device=dev)) for _ in range(n)]
rf_mb = 4.0 * (128 * 416 + 6 * 384 * 1248 + 7 * 384 * 1248) / 1e6
report("refine front end (fused)", timeit(lambda i: ops.refine_frontend(lows[i], *imgs[i]), n), rf_mb)
report("refine front end (torch ops, reference arithmetic)", timeit(lambda i: refine_frontend_torch(lows[i], *imgs[i]), n), rf_mb)
if "--json" in sys.argv:
    print(json.dumps(results))
