"""Kernel timeline of one hot-path step (CUPTI through torch.profiler; there is no nsys in the image).

    python profiles/timeline.py [--graph] [--batch B] > profiles/r02/timeline.txt

Prints every kernel of the LAST profiled step with its stream, start offset and duration, plus the idle time of the
device (no kernel resident) and the per-kernel-name totals.  With --graph the step is a CUDA-graph replay (what
bench.py times); without it, eager launches.  Concurrent kernels on different streams show as overlapping rows --
this is the view ncu (which serialises launches) cannot give.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--graph", action="store_true")
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--all", action="store_true", help="include memset / memcpy activities")
args = ap.parse_args()
dev = torch.device("cuda:0")
hp = bench.make_hot_path().to(dev)
sets = bench.make_inputs(args.batch, 3, dev)
with torch.no_grad():
    for s in sets:
        hp(*s)
    torch.cuda.synchronize()
    if args.graph:
        graphs = []
        for s in sets:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                hp(*s)
            graphs.append(g)
        step = lambda i: graphs[i % len(graphs)].replay()
    else:
        step = lambda i: hp(*sets[i % len(sets)])
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(args.steps):
            step(i)
            torch.cuda.synchronize()

ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None
      and (args.all or ("memcpy" not in e.name.lower() and "memset" not in e.name.lower()))]
ev.sort(key=lambda e: e.time_range.start)
# split into steps at gaps > 200 us and keep the last one
steps, cur = [], [ev[0]]
for a, b in zip(ev, ev[1:]):
    if b.time_range.start - a.time_range.end > 200:
        steps.append(cur)
        cur = []
    cur.append(b)
steps.append(cur)
last = steps[-1]
t0 = last[0].time_range.start
span = max(e.time_range.end for e in last) - t0
print("# %d kernels, span %.1f us (%s, B=%d)" % (len(last), span, "graph replay" if args.graph else "eager", args.batch))
streams = sorted({e.device_resource_id if hasattr(e, "device_resource_id") else 0 for e in last})
print("# %-10s %9s %8s  %s" % ("stream", "start_us", "dur_us", "kernel"))
busy_until, idle = t0, 0.0
tot = {}
for e in last:
    s = getattr(e, "device_resource_id", 0)
    st, en = e.time_range.start, e.time_range.end
    if st > busy_until:
        idle += st - busy_until
    busy_until = max(busy_until, en)
    name = e.name[:90]
    tot.setdefault(name, [0, 0.0])
    tot[name][0] += 1
    tot[name][1] += en - st
    print("  %-10s %9.1f %8.1f  %s" % (streams.index(s), st - t0, en - st, name))
print("# device idle (no kernel resident) %.1f us of %.1f" % (idle, span))
print("# totals by kernel")
for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("#  %4d x %8.1f us  %s" % (n, t, name))
