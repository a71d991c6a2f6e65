"""A/B of the halo engine against the gather engine on the dense layers of config 2 (1/3 scale), CUDA events over
a captured graph of calls rotating through inputs larger than L2.  AANET_HALO is read per launch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from aanet_b200 import ops  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    C, H, W, n = 64, 128, 416, 8
    xs = [torch.randn(1, H, W, C, device=dev) for _ in range(n)]
    sc, sh = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
    layers = {
        "3x3 64->64 dil1 (SimpleBottleneck conv2)": (torch.randn(C, C, 3, 3, device=dev) / 24, 3, 1, 1, 1, C),
        "1x1 64->64 (conv1/conv3)": (torch.randn(C, C, 1, 1, device=dev) / 8, 1, 0, 1, 1, C),
        "3x3 64->54 dil2 groups2 (offset head)": (torch.randn(54, 32, 3, 3, device=dev) / 17, 3, 2, 2, 2, 54),
    }
    for name, (w, k, pad, dil, grp, Co) in layers.items():
        wp = ops.pack_conv_weight(w, grp)
        head = Co == 54
        scv, shv = (None, None) if head else (sc, sh)

        def call(i):
            return ops.conv2d_nhwc(xs[i], wp, Co, k, k, None, scv, shv, None,
                                   ops.ACT_OFFSET_MASK if head else ops.ACT_RELU, 0.0, 1, pad, dil, grp,
                                   out_nchw=head, n_offset_ch=36 if head else 0, mask_scale=2.0)
        res = {}
        for flag in ("0", "1"):
            os.environ["AANET_HALO"] = flag
            res[flag] = bench._timed(call, n, 24, dev) * 1e3
        os.environ["AANET_HALO"] = "0"
        os.environ["AANET_DENSE_TMEM"] = "1"
        for g in ("3", "4"):
            os.environ["AANET_DENSE_GROUPS"] = g
            print("   TMEM-A dense kernel, %s producer groups: %.1f us" % (g, bench._timed(call, n, 24, dev) * 1e3))
        del os.environ["AANET_DENSE_GROUPS"]
        os.environ["AANET_DENSE_TMEM"] = "0"
        os.environ["AANET_HALO"] = "1"
        os.environ["AANET_HALO_ROT"] = "0"
        norot = bench._timed(call, n, 24, dev) * 1e3
        del os.environ["AANET_HALO_ROT"]
        print("   without the per-CTA tap rotation: %.1f us" % norot)
        sweep = []
        for slots in (2, 3):
            for bst in (2, 4, 8):
                os.environ["AANET_HALO_SLOTS"], os.environ["AANET_HALO_BST"] = str(slots), str(bst)
                sweep.append("s%d/b%d %.1f" % (slots, bst, bench._timed(call, n, 24, dev) * 1e3))
        del os.environ["AANET_HALO_SLOTS"], os.environ["AANET_HALO_BST"]
        print("   sweep (halo slots / weight stages, us):", "  ".join(sweep))
        os.environ["AANET_HALO"] = "1"
        a = call(0)
        os.environ["AANET_HALO"] = "0"
        b = call(0)
        print("%-44s gather %6.1f us   halo %6.1f us   max|diff| %.2e" % (name, res["0"], res["1"],
                                                                          float((a - b).abs().max())))


if __name__ == "__main__":
    main()
