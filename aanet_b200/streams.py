"""Fork/join of independent per-scale work onto side streams.

The three pyramid scales of the hot path are independent inside the cost-volume stage and inside the ISA
half of every aggregation module.  The 1/6 and 1/12 scale kernels are launch/latency bound (26-104 tiles
on 148 SMs), so running them next to the 1/3 scale work hides them almost completely.  The pattern is
plain event fork/join, which CUDA-graph capture records as parallel branches.

Every kernel of this path takes a whole SM per CTA (150-220 KB of shared memory, tensor memory), so a stage costs about
(sum over its kernels of CTAs x duration) / 148 however the branches are ordered: measured in round 2, enqueueing the
1/3-scale branch first (AANET_FORK_ORDER=1), giving it a high-priority stream, or running the CSA stage as a task graph
with one stream per exchange chain all came out 1-6 % SLOWER than this plain fork (1185 pairs/s) -- the hardware block
scheduler packs the small kernels into the gaps of the large ones best when they are enqueued first.

Allocator note: a tensor produced on a side stream and consumed on the main stream (or vice versa) must not
be freed while the other stream may still touch it.  Callers keep every intermediate alive until the join
that ends their forward (see FusedAggregation.__call__), so blocks only return to their pools after all
consumers were enqueued and the next fork orders any reuse behind them.
"""
import os

import torch

_side = {}


def side_streams(device, n, first=0):
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    pool = _side.setdefault(key, [])
    while len(pool) < first + n:
        pool.append(torch.cuda.Stream(device))
    return pool[first:first + n]


# A/B switch: AANET_FORK_ORDER=1 enqueues the main branch before the side branches
LEAD_FIRST = os.environ.get("AANET_FORK_ORDER", "0") == "1"


def fork_join(device, fns, first=0):
    """Run fns[0] on the current stream and fns[1:] on side streams; returns their results in order.
    first: index of the first side stream of the pool to use (a fork nested inside a branch of another fork must
    not share its side streams)."""
    if len(fns) == 1:
        return [fns[0]()]
    main = torch.cuda.current_stream(device)
    start = torch.cuda.Event()
    start.record(main)
    results = [None] * len(fns)
    done = []
    if LEAD_FIRST:
        results[0] = fns[0]()
    for i, (fn, s) in enumerate(zip(fns[1:], side_streams(device, len(fns) - 1, first)), 1):
        s.wait_event(start)
        with torch.cuda.stream(s):
            results[i] = fn()
            e = torch.cuda.Event()
            e.record(s)
        done.append(e)
    if not LEAD_FIRST:
        results[0] = fns[0]()
    for e in done:
        main.wait_event(e)
    return results
