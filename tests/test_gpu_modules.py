"""GPU tests of the drop-in modules (reference API) against golden outputs of the reference's own
modules and against the CPU port (oracle/torch_port.py) at larger sizes."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import torch_port as port

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_convs():
    """The parity bar (1e-4) needs true fp32 in the cuDNN glue convs (SURVEY.md section 7 pitfalls)."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _sd(z, pre):
    return {k[len(pre):]: torch.from_numpy(v) for k, v in z.items() if k.startswith(pre)}


def npy(t):
    return t.detach().float().cpu().numpy()


def test_deform_conv2d_golden(golden):
    import aanet_b200.nets as n
    z = golden("deform_layer")
    layer = n.DeformConv2d(8, 8, dilation=2, deformable_groups=2)
    layer.load_state_dict(_sd(z, "dc2d_sd/"), strict=True)
    layer.cuda()
    out = layer(torch.from_numpy(z["dc2d_x"]).cuda())
    assert rel_err(npy(out), z["dc2d_out"]) < 1e-4


@pytest.mark.parametrize("grad", [False, True])
def test_deform_bottleneck_golden(golden, grad):
    import aanet_b200.nets as n
    z = golden("deform_layer")
    blk = n.DeformSimpleBottleneck(8, 8, mdconv_dilation=2, deformable_groups=2)
    blk.load_state_dict(_sd(z, "blk_sd/"), strict=True)
    blk.cuda().eval()
    x = torch.from_numpy(z["blk_x"]).cuda()
    with torch.set_grad_enabled(grad):        # grad=False takes the fused BN+ReLU epilogue
        out = blk(x)
    assert rel_err(npy(out), z["blk_out"]) < 1e-4


@pytest.mark.parametrize("name", ["agg", "agg_inter", "agg32"])
@pytest.mark.parametrize("grad", [False, True])
def test_hot_path_golden(golden, name, grad):
    """CostVolumePyramid -> AdaptiveAggregation -> DisparityEstimation vs the reference's outputs.
    agg32 (D0 = 32) with grad=False is the reference-made fixture that runs on the fused tcgen05 executor."""
    import aanet_b200.nets as n
    z = golden(name)
    D0, inter = int(z["D0"]), bool(z["inter"])
    agg = n.AdaptiveAggregation(D0, num_scales=3, num_fusions=6, num_stage_blocks=1, num_deform_blocks=3,
                                intermediate_supervision=inter, deformable_groups=2, mdconv_dilation=2)
    missing, unexpected = agg.load_state_dict(_sd(z, "sd/"), strict=False)
    assert not unexpected and all(k.endswith("num_batches_tracked") for k in missing)
    agg.cuda().eval()
    Ls = [torch.from_numpy(z["L%d" % s]).cuda() for s in range(3)]
    Rs = [torch.from_numpy(z["R%d" % s]).cuda() for s in range(3)]
    with torch.set_grad_enabled(grad):
        costs = n.CostVolumePyramid(D0)(Ls, Rs)
        for s in range(3):
            assert rel_err(npy(costs[s]), z["cost%d" % s]) < 1e-4
        outs = agg(costs)
        est = n.DisparityEstimation(D0)
        disps = [est(o) for o in reversed(outs)]
    assert len(outs) == (3 if inter else 1)
    if name == "agg32":
        assert hasattr(agg, "_aanet_fused") == (not grad)      # the fused executor ran iff autograd was off
    for i, o in enumerate(outs):
        assert rel_err(npy(o), z["agg%d" % i]) < 1e-4
    for i, d in enumerate(disps):
        assert np.abs(npy(d) - z["disp%d" % i]).max() < 1e-3


def test_hot_path_vs_port_midsize():
    """Random-init AANet-shaped aggregation (D0=32, 48x96 at the 1/3 scale) against the CPU port."""
    import aanet_b200.nets as n
    torch.manual_seed(326)
    D0, H, W, C = 32, 48, 96, 32
    agg = n.AdaptiveAggregation(D0, num_deform_blocks=3, intermediate_supervision=False).eval()
    for name, m in agg.named_modules():       # non-trivial offsets / BN statistics
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.8, 1.2)
        if name.endswith("offset_conv"):
            torch.nn.init.normal_(m.weight, std=0.05); torch.nn.init.normal_(m.bias, std=0.5)
    Ls = [torch.relu(torch.randn(2, C, H >> s, W >> s)) for s in range(3)]
    Rs = [torch.relu(torch.randn(2, C, H >> s, W >> s)) for s in range(3)]
    sd = {k: v.clone() for k, v in agg.state_dict().items()}
    ref = port.hot_path(Ls, Rs, sd, D0, corr_c=True)
    agg.cuda()
    with torch.no_grad():
        costs = n.CostVolumePyramid(D0)([t.cuda() for t in Ls], [t.cuda() for t in Rs])
        outs = agg(costs)
        disp = n.DisparityEstimation(D0)(outs[0])
    assert np.abs(npy(disp) - ref[-1].numpy()).max() < 1e-3


@pytest.mark.parametrize("D0,H,W,B", [(32, 24, 40, 2), (64, 128, 416, 1)])
def test_fused_executor_matches_module_path(D0, H, W, B):
    """Channels-last tcgen05 executor (eval + no_grad) vs the module-by-module path on the same weights,
    at a small size and at the full KITTI 1/3-scale pyramid of the bench workload."""
    import aanet_b200.nets as n
    from aanet_b200 import fused
    torch.manual_seed(326)
    agg = n.AdaptiveAggregation(D0, num_deform_blocks=3, intermediate_supervision=False).cuda().eval()
    for name, m in agg.named_modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.8, 1.2)
        if name.endswith("offset_conv"):
            torch.nn.init.normal_(m.weight, std=0.05); torch.nn.init.normal_(m.bias, std=0.5)
    assert fused.supported(agg)
    costs = [torch.randn(B, D0 >> s, H >> s, W >> s, device="cuda") for s in range(3)]
    with torch.no_grad():
        fast = agg([c.clone() for c in costs])
        agg.use_fused_inference = False
        slow = agg([c.clone() for c in costs])
        d_fast = n.DisparityEstimation(D0)(fast[0])
        d_slow = n.DisparityEstimation(D0)(slow[0])
    assert hasattr(agg, "_aanet_fused")
    assert rel_err(npy(fast[0]), npy(slow[0])) < 1e-4
    assert (d_fast - d_slow).abs().max().item() < 1e-3
    # editing a weight invalidates the packed cache
    with torch.no_grad():
        agg.final_conv[0].bias.add_(1.0)
        agg.use_fused_inference = True
        again = agg([c.clone() for c in costs])
    assert rel_err(npy(again[0]), npy(slow[0] + 1.0)) < 1e-4


def _bench_init(hp, randomize):
    """bench.py's weight init (offset_conv ~ N(0, 0.05^2)); `randomize` additionally gives the BatchNorms
    non-trivial statistics and the offset head a bias of +-0.5 px so that samples leave the image."""
    for nm, m in hp.named_modules():
        if nm.endswith("offset_conv"):
            torch.nn.init.normal_(m.weight, std=0.05)
            torch.nn.init.normal_(m.bias, std=0.5 if randomize else 0.05)
        if randomize and isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.8, 1.2)
            torch.nn.init.uniform_(m.weight, 0.8, 1.2); torch.nn.init.normal_(m.bias, std=0.1)


@pytest.mark.parametrize("name,D0,Cs,H,W,B,randomize", [
    ("config2 AANet KITTI 384x1248 (the bench workload, bench init)", 64, (128, 128, 128), 128, 416, 1, False),
    ("config2, randomised BN statistics and offset bias", 64, (128, 128, 128), 128, 416, 1, True),
    ("config3 AANet+ Scene Flow 576x960 (2 of the 64 pairs)", 64, (32, 64, 128), 192, 320, 2, False),
    ("config5 AANet 1080p -> 1104x1920, max_disp 288", 96, (128, 128, 128), 368, 640, 1, False),
])
def test_hot_path_full_config_vs_port(name, D0, Cs, H, W, B, randomize):
    """The benchmarked configurations at FULL size against the CPU port of the reference
    (oracle/torch_port.py: torch CPU convs + torchvision deform_conv2d + the C oracle's correlation / CSA /
    soft-argmin) -- not against this repo's own module path: cost volumes and the aggregated volume within 1e-4
    relative, disparity within 1e-3 px (BASELINE north_star), through HotPath exactly as bench.py drives it
    (fused channels-last executor: the MODE 2 deformable engine at D0 = 64, two N tiles per layer at D0 = 96).
    Mirrors nets/aanet.py:216-219."""
    import aanet_b200.nets as n
    from aanet_b200 import fused
    from aanet_b200.pipeline import HotPath
    torch.manual_seed(326)
    hp = HotPath(3 * D0, num_deform_blocks=3, intermediate_supervision=False).eval()
    _bench_init(hp, randomize)
    g = torch.Generator().manual_seed(326)
    Ls = [torch.relu(torch.randn(B, Cs[s], H >> s, W >> s, generator=g)) for s in range(3)]
    Rs = [torch.relu(torch.randn(B, Cs[s], H >> s, W >> s, generator=g)) for s in range(3)]
    sd = {k: v.clone() for k, v in hp.aggregation.state_dict().items()}
    try:
        import torchvision  # noqa: F401
        impl = "tv"
    except Exception:
        impl = "c"
    torch.set_num_threads(max(1, (os.cpu_count() or 2)))
    with torch.no_grad():
        ref_costs = port.cost_volume_pyramid(Ls, Rs, D0, use_c=True)
        ref_agg = port.adaptive_aggregation(ref_costs, sd, impl=impl)
        ref_disp = port.disparity_estimation(ref_agg[0], True)
    hp.cuda()
    assert fused.supported(hp.aggregation)
    Lc, Rc = [t.cuda() for t in Ls], [t.cuda() for t in Rs]
    with torch.no_grad():
        costs = n.CostVolumePyramid(D0)(Lc, Rc)
        agg = hp.aggregation([c.clone() for c in costs])
        disp = hp(Lc, Rc)[-1]               # correlation_nhwc -> fused executor -> soft-argmin (the bench's path)
    assert hasattr(hp.aggregation, "_aanet_fused")
    for s in range(3):
        assert rel_err(npy(costs[s]), ref_costs[s].numpy()) < 1e-4, "cost volume, scale %d" % s
    assert rel_err(npy(agg[0]), ref_agg[0].numpy()) < 1e-4
    assert np.abs(npy(disp) - ref_disp.numpy()).max() < 1e-3
    assert np.abs(npy(n.DisparityEstimation(D0)(agg[0])) - ref_disp.numpy()).max() < 1e-3


@pytest.mark.parametrize("name,D0,C,H,W,B", [
    ("config3 AANet+ Scene Flow 576x960 (2 of the 64 pairs)", 64, 32, 192, 320, 2),
    ("config5 1080p -> 1104x1920, max_disp 288", 96, 16, 368, 640, 1),
])
def test_hot_path_large_configs(name, D0, C, H, W, B):
    """BASELINE configs 3 and 5 as parity cases: the fused tcgen05 executor against the module-by-module path
    (cuDNN fp32 glue + the same operators) at full resolution, plus size-independent properties of the cost
    volume (zero band, d = 0 plane) and of the disparity (range).  D0 = 96 exercises two N tiles per layer."""
    import aanet_b200.nets as n
    torch.manual_seed(326)
    agg = n.AdaptiveAggregation(D0, num_deform_blocks=3, intermediate_supervision=False).cuda().eval()
    for nm, m in agg.named_modules():
        if nm.endswith("offset_conv"):
            torch.nn.init.normal_(m.weight, std=0.05); torch.nn.init.normal_(m.bias, std=0.05)
    Ls = [torch.relu(torch.randn(B, C << s if D0 == 64 else 128, H >> s, W >> s, device="cuda")) for s in range(3)]
    Rs = [torch.relu(torch.randn_like(l)) for l in Ls]
    with torch.no_grad():
        costs = n.CostVolumePyramid(D0)(Ls, Rs)
        for s, c in enumerate(costs):
            Ds = D0 >> s
            assert c.shape == (B, Ds, H >> s, W >> s)
            assert torch.all(c[:, Ds - 1, :, :Ds - 1] == 0)
            assert rel_err(npy(c[:, 0]), npy((Ls[s] * Rs[s]).mean(1))) < 1e-5
        fast = agg([c.clone() for c in costs])
        agg.use_fused_inference = False
        slow = agg([c.clone() for c in costs])
        d_fast, d_slow = n.DisparityEstimation(D0)(fast[0]), n.DisparityEstimation(D0)(slow[0])
    assert rel_err(npy(fast[0]), npy(slow[0])) < 1e-4
    assert (d_fast - d_slow).abs().max().item() < 1e-3
    assert float(d_fast.min()) >= 0 and float(d_fast.max()) <= D0 - 1


def test_bf16_cost_volume_epe():
    """BASELINE config 5: bf16 cost-volume variant.  End-point error (metric.py:7-14 = mean |d_a - d_b|) of the
    1/3-scale disparity against the fp32 path, same random-init aggregation weights.  Stated tolerance:
    EPE <= 0.005 px (disparity range 0..95); measured on the B200: 0.00015 px, max |diff| 0.0009 px."""
    import aanet_b200.nets as n
    torch.manual_seed(326)
    D0, H, W = 96, 184, 320          # half of the 1104x1920 1/3-scale plane, keeps the test quick
    agg = n.AdaptiveAggregation(D0, num_deform_blocks=3, intermediate_supervision=False).cuda().eval()
    for nm, m in agg.named_modules():
        if nm.endswith("offset_conv"):
            torch.nn.init.normal_(m.weight, std=0.05); torch.nn.init.normal_(m.bias, std=0.05)
    Ls = [torch.relu(torch.randn(1, 128, H >> s, W >> s, device="cuda")) for s in range(3)]
    Rs = [torch.relu(torch.randn_like(l)) for l in Ls]
    cv, est = n.CostVolumePyramid(D0), n.DisparityEstimation(D0)
    with torch.no_grad():
        d32 = est(agg(cv(Ls, Rs))[0])
        d16 = est(agg(cv([l.bfloat16() for l in Ls], [r.bfloat16() for r in Rs]))[0])
    epe = (d32 - d16).abs().mean().item()
    print("bf16 cost volume: EPE vs fp32 = %.5f px, max |diff| = %.4f px" % (epe, (d32 - d16).abs().max().item()))
    assert epe <= 0.005


def test_training_step_runs():
    """fwd + bwd through the drop-in modules in train mode (BN batch statistics, autograd kernels)."""
    import aanet_b200.nets as n
    torch.manual_seed(0)
    agg = n.AdaptiveAggregation(16, num_deform_blocks=3, intermediate_supervision=True).cuda().train()
    for name, m in agg.named_modules():
        if name.endswith("offset_conv"):
            torch.nn.init.normal_(m.weight, std=0.05)
    Ls = [torch.relu(torch.randn(2, 8, 24 >> s, 36 >> s, device="cuda")).requires_grad_() for s in range(3)]
    Rs = [torch.relu(torch.randn(2, 8, 24 >> s, 36 >> s, device="cuda")) for s in range(3)]
    outs = agg(n.CostVolumePyramid(16)(Ls, Rs))
    loss = sum(n.DisparityEstimation(16)(o).mean() for o in outs)
    loss.backward()
    assert all(torch.isfinite(l.grad).all() for l in Ls)
    g = agg.fusions[3].branches[0][0].conv2.offset_conv.weight.grad
    assert g is not None and torch.isfinite(g).all() and g.abs().sum() > 0


def test_host_pipeline_matches_direct_forward():
    """HostPipeline (pinned host buffers in, pinned disparity out, graph replay on two slots) returns what a
    direct forward of the same HotPath returns -- through the one-DMA staging block and through per-tensor
    copies, with the slots reused several times and a different pair in every submission."""
    from aanet_b200.pipeline import HostPipeline, HotPath
    torch.manual_seed(326)
    dev = torch.device("cuda:0")
    hp = HotPath(96, num_deform_blocks=3, intermediate_supervision=False).to(dev).eval()
    shapes = [(1, 32, 24 >> s, 48 >> s) for s in range(3)]
    pairs = [([torch.relu(torch.randn(s)) for s in shapes], [torch.relu(torch.randn(s)) for s in shapes])
             for _ in range(5)]
    with torch.no_grad():
        want = [hp([t.to(dev) for t in L], [t.to(dev) for t in R])[-1].cpu() for L, R in pairs]
    pipe = HostPipeline(hp, shapes, dev)
    assert pipe.h2d_bytes == 2 * 4 * sum(int(np.prod(s)) for s in shapes)
    got = []
    for k, (L, R) in enumerate(pairs):
        if k % 2 == 0:                         # contiguous staging block, one copy
            hL, hR = pipe.staging()
            for dst, src in zip(hL + hR, L + R):
                dst.copy_(src)
            slot = pipe.submit()
        else:                                  # caller-owned pinned tensors, one copy per tensor
            slot = pipe.submit([t.pin_memory() for t in L], [t.pin_memory() for t in R])
        got.append(HostPipeline.result(slot).clone())
    for g, w in zip(got, want):
        assert torch.equal(g, w)


def test_stereodrnet_refinement_golden(golden):
    """Whole StereoDRNetRefinement (fused front end + cuDNN glue) vs the reference module's output, with the
    reference's state_dict loaded strict=True."""
    import aanet_b200.nets as n
    z = golden("refinement")
    net = n.StereoDRNetRefinement().eval()
    net.load_state_dict(_sd(z, "drnet_sd/"), strict=False)        # num_batches_tracked is not stored
    net.cuda()
    for tag in ("x3", "same", "odd", "x2"):
        with torch.no_grad():
            out = net(torch.from_numpy(z[tag + "_low"]).cuda(), torch.from_numpy(z[tag + "_left"]).cuda(),
                      torch.from_numpy(z[tag + "_right"]).cuda())
        ref = z[tag + "_drnet_out"]
        assert np.abs(npy(out) - ref).max() < 1e-3 * max(1.0, np.abs(ref).max())     # disparity, px


def test_hourglass_refinement_runs_and_trains():
    """HourglassRefinement: inference through the fused front end + the tcgen05 deformable layers, and one
    backward through the differentiable front end; both paths agree."""
    import aanet_b200.nets as n
    torch.manual_seed(326)
    net = n.HourglassRefinement().cuda().eval()
    low = (torch.rand(1, 16, 24, device="cuda") * 8)
    left, right = torch.rand(1, 3, 32, 48, device="cuda"), torch.rand(1, 3, 32, 48, device="cuda")
    with torch.no_grad():
        a = net(low, left, right)
    lg = low.clone().requires_grad_()
    b = net(lg, left, right)
    assert a.shape == (1, 32, 48) and torch.isfinite(a).all()
    assert (a - b.detach()).abs().max() < 1e-3
    b.sum().backward()
    assert lg.grad is not None and torch.isfinite(lg.grad).all() and lg.grad.abs().sum() > 0


def test_hot_path_batch_slicing():
    """A batch that does not fit one pass is run in slices; the result equals the single pass."""
    from aanet_b200.pipeline import HotPath
    torch.manual_seed(326)
    hp = HotPath(48, num_deform_blocks=3).cuda().eval()
    L = [torch.relu(torch.randn(5, 16, 24 >> s, 40 >> s, device="cuda")) for s in range(3)]
    R = [torch.relu(torch.randn(5, 16, 24 >> s, 40 >> s, device="cuda")) for s in range(3)]
    with torch.no_grad():
        assert hp.pairs_per_pass(L) == 5                      # tiny volumes: everything fits
        whole = hp(L, R)
        hp.max_pairs_per_pass = 2                             # 2 + 2 + 1
        sliced = hp(L, R)
    assert len(whole) == len(sliced)
    for a, b in zip(whole, sliced):
        assert a.shape == b.shape and torch.equal(a, b)


def test_captured_graph_sees_weight_updates():
    """the fused executor runs on private packed copies of the weights whose addresses are baked into
    captured graphs.  After an in-place weight update + one eager forward (which re-packs into the SAME buffers),
    replaying the OLD graph must give the NEW result, and the pack buffers must not have moved."""
    from aanet_b200.pipeline import HotPath
    torch.manual_seed(326)
    hp = HotPath(96, num_deform_blocks=3).cuda().eval()
    L = [torch.relu(torch.randn(1, 32, 24 >> s, 48 >> s, device="cuda")) for s in range(3)]
    R = [torch.relu(torch.randn(1, 32, 24 >> s, 48 >> s, device="cuda")) for s in range(3)]
    graph, outs = hp.capture(L, R)
    graph.replay(); torch.cuda.synchronize()
    before = outs[-1].clone()
    fz = hp.aggregation._aanet_fused
    ptr = fz.final[0].wpack.data_ptr()
    with torch.no_grad():
        hp.aggregation.final_conv[0].weight.mul_(0.5)
        hp.aggregation.fusions[0].branches[0][0].conv2.weight.add_(0.01)
        want = hp(L, R)[-1].clone()                      # eager forward: re-packs in place
    assert hp.aggregation._aanet_fused is fz and fz.final[0].wpack.data_ptr() == ptr
    graph.replay(); torch.cuda.synchronize()
    assert not torch.equal(before, want)
    assert torch.equal(outs[-1], want)


def test_fused_executor_peak_memory():
    """The executor's lifetime plan: peak extra memory of one aggregation pass stays below 8 volumes of the
    1/3-scale size per pair (a keep-everything plan holds ~70), so configs 3 and 5 run in one pass."""
    import aanet_b200.nets as n
    torch.manual_seed(326)
    D0, H, W, B = 64, 96, 160, 4
    agg = n.AdaptiveAggregation(D0, num_deform_blocks=3, intermediate_supervision=False).cuda().eval()
    costs = [torch.randn(B, H >> s, W >> s, D0 >> s, device="cuda") for s in range(3)]   # channels-last
    from aanet_b200 import fused
    with torch.no_grad():
        fused.run(agg, costs, nhwc=True)                 # packs the weights
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        base = torch.cuda.memory_allocated()
        out = fused.run(agg, costs, nhwc=True)
        torch.cuda.synchronize()
    vol = 4 * B * D0 * H * W
    peak = (torch.cuda.max_memory_allocated() - base) / vol
    print("peak extra memory: %.2f volumes of the 1/3-scale size" % peak)
    assert out[0].shape == (B, D0, H, W)
    assert peak < 8.0


def test_fused_softargmin_epilogue_matches_separate_kernels():
    """HotPath with the final 1x1 convolution + soft-argmin in one launch (ops.ACT_SOFTARGMIN) against the final
    convolution followed by the soft-argmin kernel: <= 1e-4 px (both are fp32 online softmaxes), and both against the
    CPU port within the path's 1e-3 px."""
    import os
    from aanet_b200.pipeline import HotPath
    torch.manual_seed(5)
    hp = HotPath(96, num_deform_blocks=1, num_fusions=2).cuda().eval()
    for name, m in hp.named_modules():
        if name.endswith("offset_conv"):
            torch.nn.init.normal_(m.weight, std=0.05)
    L = [torch.relu(torch.randn(2, 32, 48 // 2 ** s, 96 // 2 ** s, device="cuda")) for s in range(3)]
    R = [torch.relu(torch.randn(2, 32, 48 // 2 ** s, 96 // 2 ** s, device="cuda")) for s in range(3)]
    old = os.environ.get("AANET_FUSE_SOFTARGMIN")
    try:
        os.environ["AANET_FUSE_SOFTARGMIN"] = "1"
        with torch.no_grad():
            fused_d = hp(L, R)
        os.environ["AANET_FUSE_SOFTARGMIN"] = "0"
        with torch.no_grad():
            two = hp(L, R)
    finally:
        if old is None:
            os.environ.pop("AANET_FUSE_SOFTARGMIN", None)
        else:
            os.environ["AANET_FUSE_SOFTARGMIN"] = old
    assert len(fused_d) == len(two) == 1 and fused_d[0].shape == two[0].shape == (2, 48, 96)
    assert float((fused_d[0] - two[0]).abs().max()) < 1e-4


def test_csa_conv1_executor_path_matches_default():
    """The opt-in executor path that produces the CSA sum inside the next module's conv1 launch
    (AANET_CSA_CONV1=1, ops.csa_conv1_nhwc) against the default (csa_fuse + conv1): <= 1e-4 px."""
    import os
    from aanet_b200.pipeline import HotPath
    torch.manual_seed(9)
    hp = HotPath(192, num_deform_blocks=1, num_fusions=3).cuda().eval()
    for name, m in hp.named_modules():
        if name.endswith("offset_conv"):
            torch.nn.init.normal_(m.weight, std=0.05)
    L = [torch.relu(torch.randn(1, 32, 40 // 2 ** s, 88 // 2 ** s, device="cuda")) for s in range(3)]
    R = [torch.relu(torch.randn(1, 32, 40 // 2 ** s, 88 // 2 ** s, device="cuda")) for s in range(3)]
    old = os.environ.get("AANET_CSA_CONV1")
    try:
        os.environ["AANET_CSA_CONV1"] = "1"
        with torch.no_grad():
            a = hp(L, R)
        os.environ["AANET_CSA_CONV1"] = "0"
        with torch.no_grad():
            b = hp(L, R)
    finally:
        if old is None:
            os.environ.pop("AANET_CSA_CONV1", None)
        else:
            os.environ["AANET_CSA_CONV1"] = old
    assert float((a[-1] - b[-1]).abs().max()) < 1e-4
