"""CPU-only tests: the C-ABI library builds, loads and exports every declared symbol; the drop-in
modules keep the reference's API / state_dict layout; host-side sharding logic (gloo, world_size 2)."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


@pytest.fixture(scope="module")
def lib_path():
    from aanet_b200 import build
    return build.build()


def test_library_exports_every_header_symbol(lib_path):
    header = open(os.path.join(ROOT, "include", "aanet_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(aanet_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 12
    lib = ctypes.CDLL(lib_path)
    for name in declared:
        assert hasattr(lib, name), "libaanet_b200.so does not export %s" % name
    from aanet_b200 import _lib
    assert set(_lib.SIGNATURES) == declared          # binding covers exactly the header
    loaded = _lib.load()
    assert loaded.aanet_abi_version() == _lib.ABI_VERSION == 2
    assert loaded.aanet_status_string(0) == b"ok"
    assert b"shape" in loaded.aanet_status_string(2)


def test_argument_errors_without_gpu(lib_path):
    """Argument validation happens before any CUDA call, so it is testable on a CPU box."""
    from aanet_b200 import _lib
    lib = _lib.load()
    assert lib.aanet_corr_fwd(None, None, None, 1, 1, 1, 1, 1, None) == 1            # NULL
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.aanet_corr_fwd(p, p, p, 1, 0, 1, 1, 1, None) == 2                      # bad shape
    assert lib.aanet_softargmin_fwd(p, p, 0, 4, 4, 4, 1, None) == 2
    # Cin not divisible by deformable groups (cpp:497-516 style check)
    assert lib.aanet_mdcn_fwd(p, p, p, p, None, p, 1, 6, 8, 8, 4, 3, 3, 1, 1, 1, 1, 4,
                              None, None, 0, None, 0, None) == 2
    assert lib.aanet_mdcn_workspace_bytes(1, 1, 64, 128, 416, 64, 3, 3, 1, 2, 2, 1, 2) > 0
    # forward workspace = tf32 hi/lo packed weights (18 K-blocks x 2 x 64 x 32 floats) + NHWC x
    assert lib.aanet_mdcn_workspace_bytes(0, 1, 64, 128, 416, 64, 3, 3, 1, 2, 2, 1, 2) == \
        18 * 2 * 64 * 32 * 4 + 64 * 128 * 416 * 4     # + channels-last copy of x
    assert lib.aanet_mdcn_workspace_bytes(0, 1, 6, 8, 8, 6, 3, 3, 1, 1, 1, 1, 1) == 0   # Cin % 4 != 0: FFMA path
    assert lib.aanet_conv2d_workspace_bytes(1, 64, 128, 416, 64, 1, 1, 1, 0, 1, 1) == \
        2 * 2 * 64 * 32 * 4 + 64 * 128 * 416 * 4
    assert lib.aanet_conv_wpack_bytes(64, 64, 3, 3, 1, 0) == 18 * 2 * 64 * 32 * 4
    assert lib.aanet_conv_wpack_bytes(54, 64, 3, 3, 2, 0) == 2 * 9 * 2 * 32 * 32 * 4     # grouped: N = 27 -> 32
    assert lib.aanet_conv_wpack_bytes(16, 16, 3, 3, 1, 64) == 5 * 2 * 64 * 32 * 4       # packed for a BN=64 batch
    assert lib.aanet_conv_wpack_bytes(128, 64, 1, 1, 1, 0) == 2 * 2 * 2 * 64 * 32 * 4   # 128 outputs = 2 N tiles
    assert lib.aanet_conv_wpack_bytes(8, 6, 3, 3, 1, 0) == 0
    assert lib.aanet_conv_batch_nhwc(None, 1, 0, 0, None) == 1
    assert lib.aanet_conv_batch_nhwc(p, 4, 0, 0, None) == 2                              # at most 3 problems
    assert lib.aanet_conv2d_fwd(p, p, None, None, None, None, 0, 0.0, p, 1, 6, 8, 8, 4, 3, 3, 1, 1, 1, 1,
                                p, 1 << 20, None) == 3                                 # unsupported Cin
    with pytest.raises(_lib.AanetError):
        _lib.check(2, "x")


def test_cpu_tensors_raise_like_the_reference():
    from aanet_b200 import ops
    import aanet_b200.nets as n
    x = torch.randn(1, 4, 6, 6)
    with pytest.raises(NotImplementedError):       # deform_conv.py:135-136
        ops.correlation(x, x, 3)
    with pytest.raises(NotImplementedError):
        n.DisparityEstimation(4)(x)
    with pytest.raises(NotImplementedError):
        n.DeformConv2d(4, 4)(x)
    with pytest.raises(NotImplementedError):
        n.CostVolume(4, 'difference')(x, x)


def test_module_api_and_state_dict_match_reference_fixture(golden):
    import aanet_b200.nets as n
    z = golden("agg_inter")
    ref_keys = {k[3:]: v.shape for k, v in z.items() if k.startswith("sd/")}
    agg = n.AdaptiveAggregation(int(z["D0"]), num_scales=3, num_fusions=6, num_stage_blocks=1,
                                num_deform_blocks=3, intermediate_supervision=True)
    mine = {k: tuple(v.shape) for k, v in agg.state_dict().items() if not k.endswith("num_batches_tracked")}
    assert set(mine) == set(ref_keys)
    for k in mine:
        assert mine[k] == tuple(ref_keys[k]), k
    # offset_conv starts at zero (deform.py:75-76); mdconv weight is U(+-1/sqrt(Cin*k*k)) (deform_conv.py:339-346)
    dc = agg.fusions[3].branches[0][0].conv2
    assert float(dc.offset_conv.weight.abs().sum()) == 0.0 and float(dc.offset_conv.bias.abs().sum()) == 0.0
    bound = 1.0 / np.sqrt(16 * 9)
    assert float(dc.deform_conv.weight.abs().max()) <= bound
    assert dc.deform_conv.bias is None
    assert dc.offset_conv.groups == 2 and dc.offset_conv.out_channels == 54
    # defaults of the public constructors
    import inspect
    sig = inspect.signature(n.DeformConv2d.__init__).parameters
    assert [sig[k].default for k in ("kernel_size", "stride", "dilation", "groups", "deformable_groups",
                                     "modulation", "double_mask", "bias")] == [3, 1, 2, 1, 2, True, True, False]
    sig = inspect.signature(n.AdaptiveAggregation.__init__).parameters
    assert sig["num_deform_blocks"].default == 2 and sig["intermediate_supervision"].default is True
    sig = inspect.signature(n.ModulatedDeformConv.__init__).parameters
    assert sig["bias"].default is True and sig["deformable_groups"].default == 1


def test_last_module_has_single_output_branch():
    import aanet_b200.nets as n
    agg = n.AdaptiveAggregation(16, num_deform_blocks=3, intermediate_supervision=False)
    assert [m.num_output_branches for m in agg.fusions] == [3, 3, 3, 3, 3, 1]
    assert len(agg.final_conv) == 1 and agg.final_conv[0].bias is not None
    kinds = [type(m.branches[0][0]).__name__ for m in agg.fusions]
    assert kinds == ["SimpleBottleneck"] * 3 + ["DeformSimpleBottleneck"] * 3


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "nets")), reason="reference checkout not present")
def test_dropin_into_unmodified_reference():
    """nets/aanet.py builds unchanged on top of the drop-in; parameter names/shapes equal the stock model's."""
    code = r"""
import sys, types, json
sys.path.insert(0, %r); sys.path.insert(0, %r)
import torch
mode = sys.argv[1]
if mode == "ours":
    import aanet_b200.dropin as d
    d.install()
    import nets
    d.patch(nets)
else:
    sys.modules["nets.deform_conv.deform_conv_cuda"] = types.ModuleType("stub")
    import nets
m = nets.AANet(192, 0, feature_type='aanet', feature_pyramid_network=True, no_intermediate_supervision=True)
out = {k: list(v.shape) for k, v in m.state_dict().items()}
# AANet+ (GANet features, hourglass refinement with deformable layers) and the StereoDRNet refinement variant
p = nets.AANet(192, 0, feature_type='ganet', feature_pyramid=True, refinement_type='hourglass',
               no_intermediate_supervision=True)
out.update({"plus/" + k: list(v.shape) for k, v in p.state_dict().items()})
out["plus/refinement_class"] = [type(r).__module__.split(".")[0] for r in p.refinement]
q = nets.AANet(192, 0, feature_type='aanet', feature_pyramid_network=True, refinement_type='stereodrnet',
               no_intermediate_supervision=True)
out.update({"drnet/" + k: list(v.shape) for k, v in q.state_dict().items() if k.startswith("refinement")})
print(json.dumps(out))
""" % (REF, ROOT)
    import json
    import subprocess
    outs = {}
    for mode in ("ours", "stock"):
        r = subprocess.run([sys.executable, "-c", code, mode], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[mode] = json.loads(r.stdout.strip().splitlines()[-1])
    assert outs["ours"].pop("plus/refinement_class") == ["aanet_b200"] * 2      # ours are really swapped in
    assert outs["stock"].pop("plus/refinement_class") == ["nets"] * 2
    assert outs["ours"] == outs["stock"]
    assert sum(int(np.prod(s)) for k, s in outs["ours"].items() if "num_batches" not in k
               and "running" not in k and "/" not in k) == 3931676
    assert sum(int(np.prod(s)) for k, s in outs["ours"].items() if k.startswith("plus/") and "num_batches" not in k
               and "running" not in k) == 8442850


# ------------------------------------------------------------------------------------ sharding (gloo)
def test_shard_range_partitions():
    from aanet_b200.sharding import shard_range
    for n in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from aanet_b200.sharding import gather_batch, max_over_ranks, shard_batch
    torch.manual_seed(0)
    full = [torch.arange(5 * 3, dtype=torch.float32).view(5, 3), torch.arange(5, dtype=torch.float32).view(5, 1)]
    mine = shard_batch(full, rank, world)
    res = mine[0] * 2 + mine[1]              # any per-sample function: no cross-sample term
    out = gather_batch(res, 5)
    mx = max_over_ranks(float(rank + 1), torch.device("cpu"))
    q.put((rank, out.tolist(), mx, [t.shape[0] for t in mine]))
    dist.destroy_process_group()


def test_batch_sharding_world2_gloo():
    world, port = 2, 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full0 = torch.arange(15, dtype=torch.float32).view(5, 3)
    full1 = torch.arange(5, dtype=torch.float32).view(5, 1)
    expect = (full0 * 2 + full1).tolist()
    sizes = {}
    for rank, out, mx, n in got:
        assert out == expect and mx == 2.0
        sizes[rank] = n[0]
    assert sizes == {0: 3, 1: 2}


def test_build_lists_every_cuda_source():
    """Every .cu file in csrc/ is compiled into the library (a forgotten file would only show up at link time
    on the GPU box) and every listed source exists."""
    from aanet_b200 import build
    have = sorted(f for f in os.listdir(build.CSRC) if f.endswith(".cu"))
    assert sorted(build.SOURCES) == have


def test_committed_bench_line_has_contract_keys():
    """The bench lines committed under profiles/ carry every key of the bench.py contract (both arms)."""
    import json
    line = json.loads(open(os.path.join(ROOT, "profiles", "bench_r01.json")).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in line, k
    assert line["config"]["workload"] and line["gpu_launches"] > 0 and line["vs_baseline"] is None
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["value"] < line["value"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(line["roofline"])
    assert abs(line["roofline"]["frac"] - line["roofline"]["achieved"] / line["roofline"]["peak"]) < 1e-9
    assert {"value", "unit", "cores", "kind", "sample"} <= set(line["cpu_baseline"])
    ref = json.loads(open(os.path.join(ROOT, "profiles", "bench_reference_arm_r01.json")).read().strip().splitlines()[-1])
    assert ref["impl"] == "reference" and ref["metric"] == line["metric"] and ref["unit"] == line["unit"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["cpu_baseline"]["value"] == ref["value"]
