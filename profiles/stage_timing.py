"""Per-stage device time of the hot path inside a CUDA graph: the path is captured as growing prefixes (everything
up to the k-th join of the three scale streams) and each prefix graph is timed; a stage's cost is the difference
between consecutive prefixes.  (Eager timing with events measures Python launch overhead, ~25 us per launch.)

    python profiles/stage_timing.py
Each aggregation module = ISA stage (three bottleneck chains in parallel) + CSA stage (exchange convs + fuse).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from aanet_b200 import ops  # noqa: E402
from aanet_b200.streams import fork_join  # noqa: E402

dev = torch.device("cuda:0")
hp = bench.make_hot_path().to(dev)
L, R = bench.make_inputs(1, 1, dev)[0]
with torch.no_grad():
    for _ in range(3):
        hp(L, R)
    agg = hp.aggregation
    fa = agg._aanet_fused
    class Stop(Exception):
        pass

    def run(upto):
        names = []
        keep = []

        def mark(name):
            names.append(name)
            if len(names) == upto:
                raise Stop()

        try:
            cost = hp.cost_volume(list(L), list(R))
            keep.extend(cost)
            mark("correlation")
            xs = fork_join(dev, [(lambda c=c: ops.nchw_to_nhwc(c)) for c in cost])
            mark("to channels-last")
            for m, (branches, fuse, slope) in enumerate(fa.stages):
                def branch(blocks, x):
                    def go():
                        y = x
                        for blk in blocks:
                            y = blk(y)
                            keep.append(y)
                        return y
                    return go
                xs = fork_join(dev, [branch(blocks, xs[s]) for s, blocks in enumerate(branches)])
                keep.extend(xs)
                mark("module %d ISA" % m)
                if fuse is None:
                    continue

                def fuse_row(row):
                    def go():
                        terms = []
                        for j, chain in enumerate(row):
                            t = xs[j]
                            for conv in chain:
                                t = conv(t)
                                keep.append(t)
                            terms.append(t)
                        return ops.csa_fuse_nhwc(terms, slope)
                    return go
                xs = fork_join(dev, [fuse_row(row) for row in fuse])
                keep.extend(xs)
                mark("module %d CSA" % m)
            outs = fork_join(dev, [(lambda s=s, conv=conv: conv(xs[s], out_nchw=True))
                                   for s, conv in enumerate(fa.final)])
            mark("final convs")
            d = [hp.disparity_estimation(a) for a in reversed(outs)]
            mark("soft-argmin")
        except Stop:
            pass
        return names, keep

    names, _ = run(10 ** 6)
    prev = 0.0
    for k in range(1, len(names) + 1):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            run(k)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            _, keep = run(k)
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) * 1e3 / 20
        print("%-22s %8.1f us   (prefix %8.1f us)" % (names[k - 1], t - prev, t))
        prev = t
