"""BASELINE config 4: ModulatedDeformConv ISA microbench -- forward + backward over channels and deformable
groups on 1/3-scale 64-disparity-sized volumes ([B,C,128,416]), against the reference's own CUDA op
(oracle/_ref, if built).  CUDA-event device time over a captured graph of calls, inputs rotated."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from aanet_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
ref_op = None
try:
    from oracle import build_ref
    ref_op = build_ref.load()
except Exception as e:
    print("reference op not loadable:", e)


def timeit(fn, n_sets, iters=8):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(max(2, n_sets)):
            fn(i % n_sets)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(i % n_sets)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


rows = []
H, W = 128, 416
cfgs = [(1, C, dg) for C in (16, 32, 64, 96, 128) for dg in (1, 2, 4, 8) if C % (4 * dg) == 0] + [(8, 64, 2)]
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    cfgs = [(1, 64, 2), (1, 32, 2), (1, 128, 2)]
for B, C, dg in cfgs:
    n = 3
    xs = [torch.randn(B, C, H, W, device=dev) for _ in range(n)]
    offs = [2 * torch.randn(B, dg * 18, H, W, device=dev) for _ in range(n)]
    ms = [2 * torch.sigmoid(torch.randn(B, dg * 9, H, W, device=dev)) for _ in range(n)]
    w = torch.randn(C, C, 3, 3, device=dev) / (C * 9) ** 0.5
    g = torch.randn(B, C, H, W, device=dev)
    fwd = timeit(lambda i: ops._mdcn_forward(xs[i], offs[i], ms[i], w, None, 1, 2, 2, 1, dg), n)
    bwd = timeit(lambda i: ops._mdcn_backward(xs[i], offs[i], ms[i], w, False, g, 1, 2, 2, 1, dg), n)
    r = {"B": B, "C": C, "dg": dg, "fwd_us": fwd, "bwd_us": bwd}
    if ref_op is not None:
        out = torch.empty(B, C, H, W, device=dev)
        e0, e1, fake = torch.empty(0, device=dev), torch.empty(0, device=dev), torch.empty(1, device=dev)
        args = (3, 3, 1, 1, 2, 2, 2, 2, 1, dg, False)
        gx, goff, gm, gw, gb = (torch.zeros_like(t) for t in (xs[0], offs[0], ms[0], w, fake))

        def ref_bwd(i):
            gx.zero_(); goff.zero_(); gm.zero_(); gw.zero_()
            ref_op.modulated_deform_conv_cuda_backward(xs[i], w, fake, e0, offs[i], ms[i], e1, gx, gw, gb, goff, gm,
                                                       g, *args)
        r["ref_fwd_us"] = timeit(lambda i: ref_op.modulated_deform_conv_cuda_forward(
            xs[i], w, fake, e0, offs[i], ms[i], out, e1, *args), n)
        r["ref_bwd_us"] = timeit(ref_bwd, n)
    rows.append(r)
    print("B=%d C=%3d dg=%d  fwd %8.1f us  bwd %9.1f us" % (B, C, dg, fwd, bwd) +
          ("   | reference fwd %8.1f  bwd %9.1f   speedup fwd %.2fx bwd %.2fx"
           % (r["ref_fwd_us"], r["ref_bwd_us"], r["ref_fwd_us"] / fwd, r["ref_bwd_us"] / bwd) if ref_op else ""),
          flush=True)
print(json.dumps(rows))
