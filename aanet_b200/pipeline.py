"""HotPath: the slice of AANet.forward this package accelerates, as one callable.

    cost pyramid (nets/aanet.py:216) -> AdaptiveAggregation (:217) -> DisparityEstimation (:156-167)

It owns no new arithmetic -- it wires the drop-in modules exactly like the reference's AANet does --
but adds what a B=1 latency-bound path needs on a B200: fp32 (non-TF32) cuDNN glue so results meet the
1e-4 parity bar, CUDA-graph capture of the whole path (hundreds of small launches per pair), and a
double-buffered host entry point (pinned H2D copy of the next pair overlapped with the current one).
"""
import contextlib

import os

import torch
import torch.nn as nn

from . import fused, ops
from .nets import AdaptiveAggregation, CostVolumePyramid, DisparityEstimation
from .nets.deform import _inference_mode
from .streams import fork_join


@contextlib.contextmanager
def exact_fp32():
    """cuDNN/cuBLAS TF32 off: the reference's convs would otherwise wobble at ~1e-3 on B200.
    These are PROCESS-GLOBAL torch flags, so the context is entered only on the module-by-module path (training,
    channel counts the engine does not take), whose glue convolutions are cuDNN's.  The fused inference executor
    never touches cuDNN/cuBLAS and leaves the flags alone; a multi-threaded caller (nn.DataParallel) of the
    module path should set the two flags once itself instead of relying on this per-call flip."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


class HotPath(nn.Module):
    def __init__(self, max_disp=192, num_scales=3, num_fusions=6, num_stage_blocks=1, num_deform_blocks=3,
                 deformable_groups=2, mdconv_dilation=2, intermediate_supervision=False):
        super().__init__()
        self.max_disp = max_disp // 3                       # aanet.py:59
        self.cost_volume = CostVolumePyramid(self.max_disp)
        self.aggregation = AdaptiveAggregation(self.max_disp, num_scales=num_scales, num_fusions=num_fusions,
                                               num_stage_blocks=num_stage_blocks,
                                               num_deform_blocks=num_deform_blocks,
                                               mdconv_dilation=mdconv_dilation,
                                               deformable_groups=deformable_groups,
                                               intermediate_supervision=intermediate_supervision)
        self.disparity_estimation = DisparityEstimation(self.max_disp, True)
        self._graphs = {}
        self.max_pairs_per_pass = None                      # None: derived from free device memory

    def _one_pass(self, left_pyramid, right_pyramid):
        agg_mod = self.aggregation
        D = [self.max_disp // 2 ** s for s in range(len(left_pyramid))]
        if (agg_mod.use_fused_inference and _inference_mode(agg_mod) and left_pyramid[0].is_cuda
                and left_pyramid[0].dtype == torch.float32 and all(d % 4 == 0 and 0 < d <= 128 for d in D)
                and fused.supported(agg_mod)):
            # fused inference: the correlation writes channels-last volumes straight into the executor's layout
            cost = fork_join(left_pyramid[0].device,
                             [(lambda l=l, r=r, d=d: ops.correlation_nhwc(l, r, d))
                              for l, r, d in zip(left_pyramid, right_pyramid, D)])
            if self.disparity_estimation.match_similarity:
                disp = fused.run(agg_mod, cost, nhwc=True, disparity=True)
                if disp is not None:
                    return list(reversed(disp))
            agg = fused.run(agg_mod, cost, nhwc=True)
        else:
            with exact_fp32():
                agg = agg_mod(self.cost_volume(list(left_pyramid), list(right_pyramid)))
        return [self.disparity_estimation(a) for a in reversed(agg)]

    def pairs_per_pass(self, left_pyramid):
        """How many pairs one inference pass may take.  The fused executor's lifetime plan (fused.py) holds the
        stage inputs plus ~3 volumes per scale: <= 8 volumes of the 1/3-scale size per pair including the cost
        volumes (measured: tests/test_gpu_modules.py::test_fused_executor_peak_memory).  Config 5 (32 pairs at
        1104x1920, D0 = 96: 90 MB per volume and pair) therefore takes ~23 GB and runs in ONE pass; slicing only
        happens when less than twice that is free."""
        if self.max_pairs_per_pass is not None:
            return max(1, int(self.max_pairs_per_pass))
        B, _, H, W = left_pyramid[0].shape
        per_pair = 8 * 4 * self.max_disp * H * W
        free, _ = torch.cuda.mem_get_info(left_pyramid[0].device)
        return max(1, min(B, int(free // 2 // max(per_pair, 1))))

    def forward(self, left_pyramid, right_pyramid):
        """Feature pyramids (finest first) -> list of disparities, coarse to fine (aanet.py:156-167)."""
        B = left_pyramid[0].shape[0]
        if torch.is_grad_enabled() or self.training or not left_pyramid[0].is_cuda:
            return self._one_pass(left_pyramid, right_pyramid)
        n = self.pairs_per_pass(left_pyramid)
        if n >= B:
            return self._one_pass(left_pyramid, right_pyramid)
        parts = [self._one_pass([t[i:i + n] for t in left_pyramid], [t[i:i + n] for t in right_pyramid])
                 for i in range(0, B, n)]
        return [torch.cat(ds, dim=0) for ds in zip(*parts)]

    # ------------------------------------------------------------------ CUDA graph
    @torch.no_grad()
    def capture(self, left_pyramid, right_pyramid, warmup=3):
        """Capture forward() on the given STATIC input tensors; returns (graph, outputs).  Replaying the
        graph recomputes `outputs` in place from whatever the input tensors hold."""
        assert not self.training
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.forward(left_pyramid, right_pyramid)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            outs = self.forward(left_pyramid, right_pyramid)
        return graph, outs


def bind_host_to_gpu_numa(device):
    """Pin the calling process to the CPUs of the NUMA node the GPU hangs off (sysfs local_cpulist of its PCI
    function), so that the pinned staging blocks allocated afterwards are node-local (on a two-socket host a
    DMA from the remote socket crosses the inter-socket link).  Returns a short description, or None when the
    topology is not visible (no sysfs, single-node hosts or VMs, restricted cpusets) -- then nothing changes."""
    try:
        pr = torch.cuda.get_device_properties(device)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        node = int(open(base + "/numa_node").read())
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if node < 0 or not use or use == allowed:
            return "numa node %d, affinity unchanged" % node
        os.sched_setaffinity(0, use)
        return "numa node %d, bound to %d of %d cpus" % (node, len(use), len(allowed))
    except (OSError, ValueError, AttributeError):
        return None


class HostPipeline:
    """Host-buffer entry point: pinned feature pyramids in, pinned disparity out, H2D/D2H inside.

    Two device slots; while slot k computes (graph replay on the compute stream), slot k^1 receives the
    next pair on the copy stream.  `submit()` enqueues one pair and returns; `result(i)` synchronises on
    that pair's event and returns its pinned disparity."""

    def __init__(self, hot_path, shapes, device, n_slots=2, dtype=torch.float32):
        """dtype: element type of the feature pyramids on the host and across PCIe.  torch.float32 is the parity
        path; torch.bfloat16 halves the bytes per pair (the cost volume is then computed from bf16 features,
        ops.correlation_bf16 -- a named NON-parity mode, its end-point error is reported by bench.py)."""
        self.hp, self.device, self.n = hot_path, device, n_slots
        # pinned blocks are allocated with the process bound to the GPU's NUMA node; the caller's CPU affinity is
        # restored afterwards (first-touch placement keeps the pages node-local)
        affinity = os.sched_getaffinity(0)
        self.numa = bind_host_to_gpu_numa(device)
        self.copy_stream = torch.cuda.Stream(device)
        self.compute_stream = torch.cuda.Stream(device)
        self.slots = []
        sizes = [int(torch.Size(s).numel()) for s in shapes] * 2          # left pyramid, then right pyramid
        total = sum(sizes)

        def views(flat):
            out, o = [], 0
            for n, shp in zip(sizes, list(shapes) * 2):
                out.append(flat[o:o + n].view(shp))
                o += n
            return out[:len(shapes)], out[len(shapes):]

        for _ in range(n_slots):
            # One flat device block and one flat pinned staging block per slot: a pair that is written into
            # `staging()` crosses PCIe as ONE DMA (six separate copies cost ~3 % of the transfer in set-up gaps).
            dev_flat = torch.zeros(total, device=device, dtype=dtype)
            host_flat = torch.zeros(total, dtype=dtype).pin_memory()
            L, R = views(dev_flat)
            hL, hR = views(host_flat)
            with torch.cuda.stream(self.compute_stream):
                graph, outs = hot_path.capture(L, R)
            torch.cuda.synchronize(device)
            host_out = torch.empty(outs[-1].shape, pin_memory=True)
            self.slots.append(dict(L=L, R=R, dev_flat=dev_flat, host_flat=host_flat, host_L=hL, host_R=hR,
                                   graph=graph, out=outs[-1], host=host_out,
                                   copied=torch.cuda.Event(), done=torch.cuda.Event(),
                                   free=torch.cuda.Event()))
            self.slots[-1]["free"].record(self.compute_stream)
        os.sched_setaffinity(0, affinity)
        self.i = 0
        self.h2d_bytes = self.slots[0]["host_flat"].element_size() * total
        self.d2h_bytes = 4 * self.slots[0]["out"].numel()

    def staging(self):
        """Pinned (left_pyramid, right_pyramid) views of the NEXT slot's contiguous staging block.  Fill them in
        place and call submit() without arguments.  The block is reused every n_slots submissions: wait for
        result() of the pair submitted n_slots calls earlier before overwriting it."""
        s = self.slots[self.i % self.n]
        return s["host_L"], s["host_R"]

    def submit(self, left_host=None, right_host=None):
        """Enqueue one pair.  Without arguments the slot's staging block (see staging()) is sent with one copy;
        with pinned host pyramids each tensor is copied separately."""
        s = self.slots[self.i % self.n]
        self.i += 1
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(s["free"])          # previous user of this slot has finished
            if left_host is None:
                s["dev_flat"].copy_(s["host_flat"], non_blocking=True)
            else:
                for dst, src in zip(s["L"] + s["R"], list(left_host) + list(right_host)):
                    dst.copy_(src, non_blocking=True)
            s["copied"].record(self.copy_stream)
        with torch.cuda.stream(self.compute_stream):
            self.compute_stream.wait_event(s["copied"])
            s["graph"].replay()
            s["host"].copy_(s["out"], non_blocking=True)
            s["done"].record(self.compute_stream)
            s["free"].record(self.compute_stream)
        return s

    @staticmethod
    def result(slot):
        slot["done"].synchronize()
        return slot["host"]
