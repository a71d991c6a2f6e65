"""A/B of the halo-staged deformable kernel against the round-1 gather engine on the bench's dominant layer
(1/3 scale of config 2, [1,64,128,416], dg 2, dil 2), for several offset magnitudes and margins."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from aanet_b200 import ops  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(326)
    C, H, W, n = 64, 128, 416, 6
    xs = [torch.randn(1, H, W, C, device=dev) for _ in range(n)]
    wp = ops.pack_conv_weight(torch.randn(C, C, 3, 3, device=dev) / 24)
    sc, sh = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
    for sigma in (0.3, 1.0, 2.0, 4.0):
        oms = [torch.cat([sigma * torch.randn(1, 36, H, W, device=dev),
                          2 * torch.sigmoid(torch.randn(1, 18, H, W, device=dev))], 1).contiguous() for _ in range(n)]

        def call(i):
            return ops.mdcn_nhwc(xs[i], oms[i], wp, C, 3, 3, None, sc, sh, True, 1, 2, 2, 1, 2, om_nchw=True)
        os.environ["AANET_DEFORM_HALO"] = "0"
        os.environ["AANET_DEFORM_TMEM"] = "0"
        old = bench._timed(call, n, 24, dev) * 1e3
        ref = call(0)
        os.environ["AANET_DEFORM_TMEM"] = "1"
        line = "offsets %.1f px: gather engine %6.1f us | tmem" % (sigma, old)
        for groups, margin in ((3, 4), (4, 4), (4, 2)):
            os.environ["AANET_DEFORM_MARGIN"] = str(margin)
            os.environ["AANET_DEFORM_GROUPS"] = str(groups)
            t = bench._timed(call, n, 24, dev) * 1e3
            err = float((call(0) - ref).abs().max())
            line += "  G%d margin<=%d: %6.1f us (max|diff| %.1e)" % (groups, margin, t, err)
        os.environ.pop("AANET_DEFORM_GROUPS", None)
        os.environ["AANET_DEFORM_TMEM"] = "0"
        os.environ["AANET_DEFORM_HALO"] = "1"
        line += " | smem-halo"
        for rows, margin in ((8, 2),):
            os.environ["AANET_DEFORM_MARGIN"] = str(margin)
            os.environ["AANET_DEFORM_ROWS"] = str(rows)
            t = bench._timed(call, n, 24, dev) * 1e3
            err = float((call(0) - ref).abs().max())
            line += "  rows %d margin %d: %6.1f us (max|diff| %.1e)" % (rows, margin, t, err)
        print(line, flush=True)
    os.environ["AANET_DEFORM_MARGIN"] = "4"
    os.environ["AANET_DEFORM_HALO"] = "0"
    os.environ["AANET_DEFORM_TMEM"] = "1"
    os.environ["AANET_HALO_PROF"] = "1"
    oms = [torch.cat([0.3 * torch.randn(1, 36, H, W, device=dev),
                      2 * torch.sigmoid(torch.randn(1, 18, H, W, device=dev))], 1).contiguous() for _ in range(1)]
    call(0)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
