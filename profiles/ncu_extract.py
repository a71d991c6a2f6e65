"""Extract the per-launch figures bench.py's `roofline` object quotes from a committed `ncu --set full` raw CSV.

    python profiles/ncu_extract.py profiles/step_r02_raw.csv "conv_umma_kernel<64, 2" [--commit HASH] [--key mdconv]

Appends/updates the entry `key` of profiles/ncu_roofline.json: dram traffic (read + write, MB per launch), tensor-pipe
activity, issue activity and duration, averaged over the launches whose kernel name matches, together with the
capture file and the commit the capture was made from.  bench.py reads this file; it never hard-codes the numbers.
"""
import argparse
import csv
import json
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "ncu_roofline.json")
TO_MB = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
COLS = {"dram_read_mb": "dram__bytes_read.sum", "dram_write_mb": "dram__bytes_write.sum",
        "duration_us": "gpu__time_duration.sum",
        "tensor_pipe_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "issue_active_pct": "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1_hit_pct": "l1tex__t_sector_hit_rate.pct", "l2_hit_pct": "lts__t_sector_hit_rate.pct",
        "registers": "launch__registers_per_thread"}


def extract(path, pattern):
    rows = list(csv.reader(open(path)))
    head, units, data = rows[0], rows[1], rows[2:]
    name_i = head.index("Kernel Name")
    hit = [r for r in data if re.search(pattern, r[name_i])]
    if not hit:
        raise SystemExit("no launch matches %r in %s" % (pattern, path))
    out = {"launches": len(hit), "kernel": hit[0][name_i][:100]}
    for key, col in COLS.items():
        if col not in head:
            out[key] = None
            continue
        i = head.index(col)
        vals = [float(r[i].replace(",", "")) for r in hit if r[i] not in ("", "n/a")]
        scale = TO_MB.get(units[i], 1.0) if "bytes" in col else ({"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[i], 1.0)
                                                                 if "duration" in col else 1.0)
        out[key] = sum(vals) / len(vals) * scale if vals else None
    out["traffic_mb"] = (out["dram_read_mb"] or 0.0) + (out["dram_write_mb"] or 0.0)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("pattern")
    ap.add_argument("--key", default="mdconv")
    ap.add_argument("--commit", default=None)
    a = ap.parse_args()
    e = extract(a.csv, a.pattern)
    e["capture"] = os.path.relpath(os.path.abspath(a.csv), os.path.dirname(HERE))
    e["pattern"] = a.pattern
    e["commit"] = a.commit or subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True,
                                             cwd=HERE).stdout.strip()
    table = json.load(open(OUT)) if os.path.exists(OUT) else {}
    table[a.key] = e
    json.dump(table, open(OUT, "w"), indent=1, sort_keys=True)
    print(json.dumps(e, indent=1))


if __name__ == "__main__":
    main()
