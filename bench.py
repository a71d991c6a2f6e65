"""bench.py -- stereo pairs/s through the AANet hot path (cost volume + ISA/CSA + soft-argmin).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

Workload (BASELINE.json configs[1]): AANet at KITTI 384x1248, max_disp=192, B=1, fp32, random-init
weights (offset_conv re-initialised N(0,0.05^2)), synthetic features relu(randn) of the
three pyramid shapes [1,128,128,416], [1,128,64,208], [1,128,32,104].  One step = one pass of the hot
path over one batch.  N>1: launched by torchrun, one rank per GPU, every rank runs its own pairs (batch
sharding, no data-path collective) -> weak scaling; value = pairs of all ranks / max-over-ranks time.

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (CUDA-graph replay, inputs in
HBM); `e2e` = the same through HostPipeline with pinned host buffers, H2D/D2H inside the timed region;
`roofline` = the dominant kernel (mdconv forward at the 1/3 scale) timed live with CUDA events;
`cpu_baseline` = the oracle port (reference's PyTorch CPU path restated) on the host cores.
`--impl reference` times that CPU port alone, same metric and config.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

H_IMG, W_IMG, MAX_DISP, FEAT_C = 384, 1248, 192, 128
METRIC = "384x1248 stereo pairs/s (cost+ISA/CSA+soft-argmin)"
UNIT = "pairs/s"
N_SETS = 4      # rotating input sets: 4 x 71.6 MB > 126 MB L2
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture
# (profiles/), MB; None until a capture of the current kernel exists.
NCU_TRAFFIC_MB = {"mdconv": 25.5}   # profiles/r01_launches_and_engine.md: 25.49 MB read + 0.0003 MB written


def pyramid_shapes(batch):
    return [(batch, FEAT_C, H_IMG // (3 * 2 ** s), W_IMG // (3 * 2 ** s)) for s in range(3)]


def make_hot_path():
    from aanet_b200.pipeline import HotPath
    torch.manual_seed(326)
    hp = HotPath(MAX_DISP, num_deform_blocks=3, intermediate_supervision=False)
    for name, m in hp.named_modules():
        if name.endswith("offset_conv"):
            torch.nn.init.normal_(m.weight, std=0.05)
            torch.nn.init.normal_(m.bias, std=0.05)
    return hp.eval()


def make_inputs(batch, n_sets, device, pin=False, seed=326):
    g = torch.Generator().manual_seed(seed)
    sets = []
    for _ in range(n_sets):
        L = [torch.relu(torch.randn(s, generator=g)) for s in pyramid_shapes(batch)]
        R = [torch.relu(torch.randn(s, generator=g)) for s in pyramid_shapes(batch)]
        if pin:
            L, R = [t.pin_memory() for t in L], [t.pin_memory() for t in R]
        elif device is not None:
            L, R = [t.to(device) for t in L], [t.to(device) for t in R]
        sets.append((L, R))
    return sets


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": float(self.samples[0][1]) if self.samples[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_port_step(sd, L, R):
    from oracle import torch_port as port
    try:
        import torchvision  # noqa: F401
        impl = "tv"      # the reference's CPU path: torchvision stands in for the CUDA-only op
    except Exception:
        impl = "c"
    with torch.no_grad():
        return port.hot_path(L, R, sd, MAX_DISP // 3, corr_c=False, impl=impl, fuse_c=False)


def run_cpu_baseline(steps, warmup, batch=1):
    """The reference's CPU implementation of the path (oracle port) on all host cores."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    hp = make_hot_path()
    sd = {k: v for k, v in hp.aggregation.state_dict().items()}
    (L, R), = make_inputs(batch, 1, None)
    for _ in range(warmup):
        cpu_port_step(sd, L, R)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        cpu_port_step(sd, L, R)
        times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"value": batch * steps / total, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d full pair(s) per step x %d steps (+%d warm-up), torch CPU convs + torchvision "
                      "deform_conv2d, per-disparity torch loop for the cost volume" % (batch, steps, warmup),
            "ms_per_step": 1e3 * total / steps}


# ------------------------------------------------------------------------------------------- GPU arm
def _timed(fn, n_sets, iters, device):
    """ms per call, device time: `iters` calls cycling through n_sets input sets are captured into one CUDA
    graph (no host launch gaps between the kernels) and the replay is bracketed by CUDA events on the
    launching stream."""
    side = torch.cuda.Stream(device)
    side.wait_stream(torch.cuda.current_stream(device))
    with torch.cuda.stream(side), torch.no_grad():
        for i in range(max(3, n_sets)):
            fn(i % n_sets)
    torch.cuda.current_stream(device).wait_stream(side)
    torch.cuda.synchronize(device)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g), torch.no_grad():
        for i in range(iters):
            fn(i % n_sets)
    g.replay()
    torch.cuda.synchronize(device)
    stream = torch.cuda.current_stream(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    g.replay()
    e1.record(stream)
    torch.cuda.synchronize(device)
    return e0.elapsed_time(e1) / iters


def time_kernels(device, iters=20):
    """Live CUDA-event timing (launching stream, inputs rotated over sets larger than L2) of the dominant
    kernel -- the ISA modulated deformable conv at the 1/3 scale as the fused path runs it (tcgen05 engine,
    channels-last, packed weights) -- and of the other engine / memory-bound kernels for context."""
    from aanet_b200 import ops
    torch.manual_seed(326)
    C, H, W = MAX_DISP // 3, H_IMG // 3, W_IMG // 3
    n = 6      # (13.6 + 11.5 + 13.6) MB per set -> 232 MB > L2
    xs = [torch.randn(1, H, W, C, device=device) for _ in range(n)]
    oms = [torch.cat([2 * torch.randn(1, H, W, 36, device=device),
                      2 * torch.sigmoid(torch.randn(1, H, W, 18, device=device))], -1).contiguous() for _ in range(n)]
    wp3 = ops.pack_conv_weight(torch.randn(C, C, 3, 3, device=device) / 24)
    wp1 = ops.pack_conv_weight(torch.randn(C, C, 1, 1, device=device) / 8)
    sc, sh = torch.rand(C, device=device) + 0.5, torch.randn(C, device=device)
    out = {}
    ms = _timed(lambda i: ops.mdcn_nhwc(xs[i], oms[i], wp3, C, 3, 3, None, sc, sh, True, 1, 2, 2, 1, 2), n, iters, device)
    flops = 2.0 * C * C * 9 * H * W
    bytes_alg = 4.0 * (C * H * W + 27 * 2 * H * W + C * H * W) + 36.0 * C * C
    out["mdconv"] = (ms, flops, bytes_alg)
    ms = _timed(lambda i: ops.conv2d_nhwc(xs[i], wp3, C, 3, 3, None, sc, sh, None, ops.ACT_RELU, 0.0, 1, 1, 1, 1), n, iters, device)
    out["conv3x3"] = (ms, flops, 4.0 * 2 * C * H * W)
    ms = _timed(lambda i: ops.conv2d_nhwc(xs[i], wp1, C, 1, 1, None, sc, sh, None, ops.ACT_RELU, 0.0, 1, 0, 1, 1), n, iters, device)
    out["conv1x1"] = (ms, flops / 9, 4.0 * 2 * C * H * W)
    cs = [torch.randn(1, C, H, W, device=device) for _ in range(n)]
    ms = _timed(lambda i: ops.soft_argmin(cs[i], True), n, iters, device)
    out["softargmin"] = (ms, 0.0, 4.0 * H * W * (C + 1))
    Ls = [torch.relu(torch.randn(1, FEAT_C, H, W, device=device)) for _ in range(4)]
    ms = _timed(lambda i: ops.correlation(Ls[i], Ls[(i + 1) % 4], C), 4, iters, device)
    out["correlation_s0"] = (ms, 2.0 * FEAT_C * H * (W * C - C * (C - 1) / 2), 4.0 * H * W * (2 * FEAT_C + C))
    return out


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def run_gpu(args):
    import torch.distributed as dist
    from aanet_b200 import ops
    from aanet_b200.pipeline import HostPipeline
    from aanet_b200.sharding import max_over_ranks

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    B = args.batch
    hp = make_hot_path().to(device)
    sets = make_inputs(B, N_SETS, device)
    with torch.no_grad():
        hp(*sets[0])                    # first pass packs the weights (one-off launches)
        launches0 = ops.LAUNCHES
        hp(*sets[0])
        per_step_launches = ops.LAUNCHES - launches0
        graphs = [hp.capture(L, R)[0] for (L, R) in sets]
    stream = torch.cuda.current_stream(device)

    # ---- device-resident throughput
    for i in range(args.warmup):
        graphs[i % N_SETS].replay()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        graphs[i % N_SETS].replay()
    e1.record(stream)
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1), device)
    clocks = sampler.summary() if sampler else None
    value = world * B * args.steps / (ms_total / 1e3)

    # ---- end to end: pinned host buffers -> HostPipeline -> pinned disparity
    # Each pair is written into the pipeline's pinned staging block (two slots, two different pairs) before the
    # timed region; every timed step then moves its 71.6 MB host -> device and its disparity device -> host.
    pipe = HostPipeline(hp, pyramid_shapes(B), device)
    for k, (L, R) in enumerate(make_inputs(B, pipe.n, None, seed=327)):
        hL, hR = pipe.slots[k]["host_L"], pipe.slots[k]["host_R"]
        for dst, src in zip(hL + hR, L + R):
            dst.copy_(src)
    for i in range(max(args.warmup, 4) + args.steps):     # the first pass over the pinned blocks is slow (564 vs 750)
        HostPipeline.result(pipe.submit())
    # K steps take only tens of milliseconds and the PCIe rate of a (virtualised) host wanders from run to run
    # (670 .. 750 pairs/s seen back to back on one box), so the K-step measurement is repeated and the median kept.
    e2e_runs = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        pending = []
        for i in range(args.steps):
            pending.append(pipe.submit())
            if len(pending) >= pipe.n:
                HostPipeline.result(pending.pop(0))
        for s in pending:
            HostPipeline.result(s)
        torch.cuda.synchronize(device)
        e2e_runs.append(max_over_ranks(time.perf_counter() - t0, device))
    e2e_s = statistics.median(e2e_runs)
    e2e = {"value": world * B * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes,
           "d2h_bytes_per_step": pipe.d2h_bytes,
           "h2d_gbs_per_gpu": B * args.steps / e2e_s * pipe.h2d_bytes / 1e9, "host_numa": pipe.numa,
           "runs_pairs_per_s": [world * B * args.steps / t for t in e2e_runs], "stat": "median of 3 x K steps"}

    out = None
    if rank == 0:
        hbm, bf16, peak_src = load_peaks()
        kt = time_kernels(device)
        k_ms, k_flops, k_bytes = kt["mdconv"]
        tf32_peak = bf16 / 2.0
        achieved = k_flops / (k_ms * 1e-3) / 1e12
        roofline = {"kernel": "conv_umma_kernel<64,DEFORM> = ISA modulated deformable conv, 1/3 scale "
                              "[1,64,128,416], dg=2, dil=2, 3xTF32 tcgen05",
                    "bound": "tensor", "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s",
                    "frac": achieved / tf32_peak, "traffic": NCU_TRAFFIC_MB.get("mdconv"),
                    "us_per_launch": k_ms * 1e3, "algorithmic_mb": k_bytes / 1e6,
                    "hbm_frac_if_memory_bound": k_bytes / (k_ms * 1e-3) / 1e9 / hbm,
                    "peak_source": peak_src + "; TF32 dense taken as bf16_tflops/2 (burst, kernel timed alone); "
                                   "logical FLOPs counted once although 3 MMAs are issued per product",
                    "other_kernels": {k: {"us": v[0] * 1e3, "tflops": v[1] / (v[0] * 1e-3) / 1e12,
                                          "gbs": v[2] / (v[0] * 1e-3) / 1e9, "hbm_frac": v[2] / (v[0] * 1e-3) / 1e9 / hbm}
                                      for k, v in kt.items() if k != "mdconv"}}
        cpu = None if args.no_cpu_baseline else run_cpu_baseline(3, 1)
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": "AANet hot path, KITTI 384x1248, max_disp=192, B=%d per GPU per step, fp32, "
                                      "random-init weights (configs[1])" % B,
                          "features": "relu(randn) [B,128,128,416],[B,128,64,208],[B,128,32,104]",
                          "parallelism": "batch-sharded x%d, no collective" % world,
                          "l2": "inputs rotated over %d sets (%.0f MB > 126 MB L2)" % (N_SETS, N_SETS * 71.6 * B),
                          "launch": "CUDA graph replay"},
               "e2e": e2e, "gpu_launches": per_step_launches * args.steps,
               "gpu_launches_per_step": per_step_launches, "clocks": clocks, "roofline": roofline,
               "cpu_baseline": cpu}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", 0))
    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = min(args.steps, 12)      # ~3.5 s of CPU per step: keep the whole run within minutes
        r = run_cpu_baseline(steps, min(args.warmup, 2), args.batch)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT,
                "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 2),
                "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "AANet hot path, KITTI 384x1248, max_disp=192, B=%d per step, fp32, "
                                       "random-init weights (configs[1])" % args.batch,
                           "note": "reference's PyTorch CPU path restated (oracle port) on the host cores"},
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback "
                         "(use --impl reference for the CPU port)")
    out = run_gpu(args)
    if out is not None:
        print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
