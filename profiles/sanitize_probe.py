"""Small, fast invocation of every kernel family for compute-sanitizer.

    compute-sanitizer --tool memcheck  python profiles/sanitize_probe.py
    compute-sanitizer --tool racecheck python profiles/sanitize_probe.py

Runs the fused executor at the three pyramid scales with the bench's channel counts (D0 = 64: engine MODE 0/1/2/3,
multi-problem launches, residual / lean / offset-mask epilogues) on a small plane, the tcgen05 correlation, the CSA
fuse, soft-argmin and the mdconv backward.  Shapes are tiny because the sanitizer slows kernels 10-100x."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from aanet_b200 import ops  # noqa: E402
from aanet_b200.pipeline import HotPath  # noqa: E402


def main():
    torch.manual_seed(326)
    dev = torch.device("cuda:0")
    hp = HotPath(192, num_deform_blocks=3).to(dev).eval()
    for name, m in hp.named_modules():
        if name.endswith("offset_conv"):
            torch.nn.init.normal_(m.weight, std=0.05)
            torch.nn.init.normal_(m.bias, std=0.5)
    H, W = 24, 52          # odd tile counts at every scale (24x52, 12x26, 6x13)
    L = [torch.relu(torch.randn(1, 128, H >> s, W >> s, device=dev)) for s in range(3)]
    R = [torch.relu(torch.randn(1, 128, H >> s, W >> s, device=dev)) for s in range(3)]
    with torch.no_grad():
        d = hp(L, R)[-1]
    torch.cuda.synchronize()
    print("fused hot path ok", float(d.mean()), ops.LAUNCHES, "launches")
    x = torch.randn(1, 64, 16, 20, device=dev, requires_grad=True)
    off = (2 * torch.randn(1, 36, 16, 20, device=dev)).requires_grad_()
    msk = (2 * torch.sigmoid(torch.randn(1, 18, 16, 20, device=dev))).requires_grad_()
    w = (torch.randn(64, 64, 3, 3, device=dev) / 24).requires_grad_()
    out = ops.modulated_deform_conv(x, off, msk, w, None, 1, 2, 2, 1, 2)
    out.sum().backward()
    torch.cuda.synchronize()
    print("mdconv fwd + bwd ok", float(x.grad.abs().mean()))
    c = ops.correlation(L[0].requires_grad_(), R[0], 64)
    ops.soft_argmin(c, True).sum().backward()
    torch.cuda.synchronize()
    print("correlation / soft-argmin fwd + bwd ok")


if __name__ == "__main__":
    main()
