"""TMEM-A kernels at the 1/3 scale of config 2: K-block hand-off trace (AANET_HALO_PROF=2), role counters (=1) and
timings under the experiment switches named on the command line.

    python profiles/tmem_trace.py [ENV=VALUE ...]        e.g.  AANET_MMA_SPIN=1
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from aanet_b200 import ops  # noqa: E402

for kv in sys.argv[1:]:
    k, v = kv.split("=")
    os.environ[k] = v
print("# switches:", " ".join(sys.argv[1:]) or "(none)")
dev = torch.device("cuda:0")
torch.manual_seed(326)
C, H, W, n = 64, 128, 416, 6
xs = [torch.randn(1, H, W, C, device=dev) for _ in range(n)]
res = [torch.randn(1, H, W, C, device=dev) for _ in range(n)]
wp3 = ops.pack_conv_weight(torch.randn(C, C, 3, 3, device=dev) / 24)
wp1 = ops.pack_conv_weight(torch.randn(C, C, 1, 1, device=dev) / 8)
wph = ops.pack_conv_weight(torch.randn(54, 32, 3, 3, device=dev) / 17, 2)
sc, sh = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
oms = [torch.cat([0.1 * torch.randn(1, 36, H, W, device=dev),
                  2 * torch.sigmoid(torch.randn(1, 18, H, W, device=dev))], 1).contiguous() for _ in range(n)]


def tail(i):
    return dict(wpack=wp1, Cout=C, scale=sc, shift=sh, residual=res[i], act=ops.ACT_RELU)


calls = {
    "dense 3x3": lambda i: ops.conv2d_nhwc(xs[i], wp3, C, 3, 3, None, sc, sh, None, ops.ACT_RELU, 0.0, 1, 1, 1, 1),
    "dense 3x3 + tail": lambda i: ops.conv2d_nhwc(xs[i], wp3, C, 3, 3, None, sc, sh, None, ops.ACT_RELU, 0.0, 1, 1, 1, 1,
                                                   tail=tail(i)),
    "offset head": lambda i: ops.conv2d_nhwc(xs[i], wph, 54, 3, 3, None, None, None, None, ops.ACT_OFFSET_MASK, 0.0, 1, 2,
                                             2, 2, out_nchw=True, n_offset_ch=36, mask_scale=2.0),
    "dcn": lambda i: ops.mdcn_nhwc(xs[i], oms[i], wp3, C, 3, 3, None, sc, sh, True, 1, 2, 2, 1, 2, om_nchw=True),
    "dcn + tail": lambda i: ops.mdcn_nhwc(xs[i], oms[i], wp3, C, 3, 3, None, sc, sh, True, 1, 2, 2, 1, 2, om_nchw=True,
                                          tail=tail(i)),
}
for name, fn in calls.items():
    print("%-18s %6.1f us" % (name, bench._timed(fn, n, 24, dev) * 1e3), flush=True)
if os.environ.get("TRACE", "1") == "1":
    for name in ("dense 3x3", "dcn"):
        for prof in ("1", "2"):
            os.environ["AANET_HALO_PROF"] = prof
            calls[name](0)
            torch.cuda.synchronize()
            print("^", name, "prof", prof, flush=True)
    os.environ["AANET_HALO_PROF"] = "0"

# the dominant launch exactly as the pipeline issues it (network's own offsets, fused tail)
if os.environ.get("PIPE", "1") == "1":
    hp = bench.make_hot_path().to(dev)
    (L, R), = bench.make_inputs(1, 1, dev)
    with torch.no_grad():
        hp(L, R)
        q = bench.capture_dominant_launch(hp, L, R, (H, W))
    if q is not None:
        sets = []
        for _ in range(n):
            qq = dict(q, x=q["x"].clone(), offmask=q["offmask"].clone())
            if q.get("tail"):
                qq["tail"] = dict(q["tail"], residual=q["tail"]["residual"].clone())
            sets.append(qq)
        print("%-18s %6.1f us   (mean |offset| %.3f px, tail %s)" % (
            "pipeline dcn", bench._timed(lambda i: ops.conv_batch([sets[i]], deform=True), n, 24, dev) * 1e3,
            float(q["offmask"][:, :36].abs().mean()), bool(q.get("tail"))), flush=True)
        q2 = dict(sets[0]); q2.pop("tail", None)
        if os.environ.get("TRACE", "1") == "1":
            for prof in ("1", "2"):
                os.environ["AANET_HALO_PROF"] = prof
                ops.conv_batch([sets[0]], deform=True)
                torch.cuda.synchronize()
                print("^ pipeline dcn prof", prof, flush=True)
            os.environ["AANET_HALO_PROF"] = "0"
