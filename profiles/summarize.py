"""Turn the committed ncu exports into the tables of r01_launches_and_engine.md.

    python profiles/summarize.py profiles/launches_r01.csv profiles/step_r01_raw.csv
"""
import collections
import csv
import io
import re
import sys


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"^aanet::", "", name)
    return re.sub(r"\(.*", "", name)[:60]


def launch_table(path):
    txt = [l for l in open(path) if l.startswith('"')]
    rows = [r for r in csv.DictReader(io.StringIO("".join(txt))) if r["Metric Name"] == "gpu__time_duration.sum"]
    first = [i for i, r in enumerate(rows) if "corr" in r["Kernel Name"]]
    last = rows[first[-3]:]                     # the last step starts with its three correlation launches
    agg = collections.OrderedDict()
    sm_time = 0.0
    for r in last:
        us = float(r["Metric Value"].replace(",", "")) / 1e3
        g = eval(r["Grid Size"])
        sm_time += us * min(g[0] * g[1] * g[2], 148) / 148
        a = agg.setdefault(short(r["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print("%d launches, %.1f us summed, %.1f us weighted by min(grid,148)/148" % (len(last), tot, sm_time))
    print("| kernel | launches | sum us | share |\n|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f %% |" % (k, a[0], a[1], 100 * a[1] / tot))


def kernel_table(path, n=24):
    rows = list(csv.reader(open(path)))
    h = rows[0]
    ix = h.index
    cols = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct",
            "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__inst_issued.avg.pct_of_peak_sustained_active"]
    print("| kernel | grid | us | DRAM read MB | DRAM write MB | L1 hit % | L2 hit % | warps active % | issue active % | regs |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for r in rows[2:2 + n]:
        g = eval(r[ix("Grid Size")])
        vals = ["%.2f" % float(r[ix(c)].replace(",", "")) for c in cols]
        print("| `%s` | %d | %s | %s |" % (short(r[ix("Kernel Name")]), g[0] * g[1] * g[2], " | ".join(vals),
                                           r[ix("launch__registers_per_thread")]))


if __name__ == "__main__":
    launch_table(sys.argv[1])
    print()
    kernel_table(sys.argv[2])
