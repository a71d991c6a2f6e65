"""Eager (no CUDA graph) hot-path steps for ncu: `python profiles/profile_step.py --steps 3`.
Every step is one pass of the hot path at the bench workload (KITTI 384x1248, B=1)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--batch", type=int, default=1)
args = ap.parse_args()
dev = torch.device("cuda:0")
hp = bench.make_hot_path().to(dev)
sets = bench.make_inputs(args.batch, 2, dev)
torch.backends.cudnn.benchmark = bool(int(os.environ.get("CUDNN_BENCHMARK", "0")))
with torch.no_grad():
    for i in range(args.steps):
        out = hp(*sets[i % 2])
        torch.cuda.synchronize()
print("ok", float(out[-1].mean()))
