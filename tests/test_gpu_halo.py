"""Halo engine (aanet_b200/csrc/halo_engine.cu): stride-1 dense convolutions whose tcgen05 A operand is a window
into a TMA-loaded, swizzled input halo -- against float64 torch convolutions, and against the gather engine."""
import os

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


FLAGS = {"window": "AANET_HALO", "tmem": "AANET_DENSE_TMEM"}


def _select(variant):
    for k in FLAGS.values():
        os.environ[k] = "0"
    if variant is not None:
        os.environ[FLAGS[variant]] = "1"


@pytest.fixture(params=["window", "tmem"])
def halo(request):
    """window: A operand = descriptor windows into the swizzled halo (halo_engine.cu); tmem: thread-per-pixel copy of
    the halo lines into tensor memory (deform_tmem.cu, DENSE mode; layers with >= 3 taps)."""
    old = {k: os.environ.get(k) for k in FLAGS.values()}
    _select(request.param)
    yield request.param
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def npy(t):
    return t.detach().float().cpu().numpy()


@pytest.mark.parametrize("cfg", [
    # B, Cin, Cout, H, W, k, pad, dil, groups
    (1, 64, 64, 128, 416, 3, 1, 1, 1),     # SimpleBottleneck conv2 at the 1/3 scale (config 2)
    (1, 64, 64, 128, 416, 1, 0, 1, 1),     # conv1 / conv3
    (1, 64, 54, 128, 416, 3, 2, 2, 2),     # offset head: grouped, dilated, 27 outputs per group
    (2, 32, 32, 33, 47, 3, 1, 1, 1),       # one channel block, ragged tiles, batch 2
    (1, 96, 96, 21, 30, 3, 2, 2, 1),       # three channel blocks, two N tiles (config 5 widths)
    (1, 128, 128, 17, 9, 3, 1, 1, 1),      # four channel blocks, two N tiles
    (1, 32, 16, 5, 7, 1, 0, 1, 1),         # image smaller than one tile
    (1, 64, 64, 20, 24, 3, 3, 1, 1),       # pad > (k-1)/2: output larger than the input
    (1, 64, 64, 20, 24, 3, 0, 1, 1),       # pad 0: output smaller than the input
])
def test_halo_conv_matches_float64(halo, cfg):
    import aanet_b200.ops as ops
    B, Ci, Co, H, W, k, pad, dil, grp = cfg
    torch.manual_seed(17)
    x = torch.randn(B, Ci, H, W, device="cuda")
    w = torch.randn(Co, Ci // grp, k, k, device="cuda") / (Ci * k * k / grp) ** 0.5
    bias = torch.randn(Co, device="cuda")
    scale, shift = torch.rand(Co, device="cuda") + 0.5, torch.randn(Co, device="cuda")
    ref = torch.nn.functional.conv2d(x.double(), w.double(), bias.double(), 1, pad, dil, grp)
    res = torch.randn_like(ref).float()
    xt, wp = ops.nchw_to_nhwc(x), ops.pack_conv_weight(w, grp)
    # residual + LeakyReLU, channels-last output (LEAN when Cout/groups % 16 == 0)
    out = ops.conv2d_nhwc(xt, wp, Co, k, k, bias, scale, shift, ops.nchw_to_nhwc(res), ops.ACT_LEAKY, 0.2, 1, pad, dil, grp)
    ref2 = torch.nn.functional.leaky_relu(ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1) + res, 0.2)
    assert rel_err(npy(out.permute(0, 3, 1, 2)), npy(ref2)) < 1e-5
    # NCHW output, ReLU, no affine
    out_nchw = ops.conv2d_nhwc(xt, wp, Co, k, k, bias, None, None, None, ops.ACT_RELU, 0.0, 1, pad, dil, grp, out_nchw=True)
    assert rel_err(npy(out_nchw), npy(torch.relu(ref))) < 1e-5
    # offset/mask head epilogue
    n_off = (Co * 2) // 3
    om = ops.conv2d_nhwc(xt, wp, Co, k, k, bias, None, None, None, ops.ACT_OFFSET_MASK, 0.0, 1, pad, dil, grp,
                         out_nchw=True, n_offset_ch=n_off, mask_scale=2.0)
    want = ref.clone()
    want[:, n_off:] = 2 * torch.sigmoid(ref[:, n_off:])
    assert rel_err(npy(om), npy(want)) < 1e-5
    # and the gather engine gives the same numbers up to the summation order of the K blocks
    _select(None)
    old = ops.conv2d_nhwc(xt, wp, Co, k, k, bias, scale, shift, ops.nchw_to_nhwc(res), ops.ACT_LEAKY, 0.2, 1, pad, dil, grp)
    _select(halo)
    assert rel_err(npy(out), npy(old)) < 1e-5


def test_halo_fused_executor_matches_gather_engine(halo):
    """The whole aggregation on the halo kernels where they apply vs the gather engine only."""
    import aanet_b200.nets as n
    torch.manual_seed(326)
    D0, H, W, B = 64, 48, 104, 2
    agg = n.AdaptiveAggregation(D0, num_deform_blocks=3, intermediate_supervision=False).cuda().eval()
    for name, m in agg.named_modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.8, 1.2)
        if name.endswith("offset_conv"):
            torch.nn.init.normal_(m.weight, std=0.05); torch.nn.init.normal_(m.bias, std=0.5)
    costs = [torch.randn(B, D0 >> s, H >> s, W >> s, device="cuda") for s in range(3)]
    with torch.no_grad():
        new = agg([c.clone() for c in costs])[0]
        _select(None)
        old = agg([c.clone() for c in costs])[0]
        _select(halo)
    assert rel_err(npy(new), npy(old)) < 1e-5


@pytest.mark.parametrize("cfg", [
    # B, C (= bottleneck width), H, W, deform, dil
    (1, 64, 128, 416, False, 1),      # SimpleBottleneck conv2 + conv3 at the 1/3 scale (config 2)
    (1, 64, 128, 416, True, 2),       # DeformSimpleBottleneck conv2 + conv3
    (2, 32, 33, 47, False, 1),        # one channel block, ragged tiles, batch 2, one K block in the tail
    (1, 64, 24, 52, True, 2),
    (3, 64, 9, 20, False, 1),         # more tiles than two per CTA never happens here; several images
    (1, 32, 24, 52, True, 2),         # the 1/6 scale: 16 channels per deformable group (two samples per K block) + tail
    (2, 32, 17, 35, True, 2),
])
def test_fused_tail_matches_two_launches(cfg):
    """conv2 (+ bn2 + ReLU) with the bottleneck's trailing conv3 + bn3 + identity + ReLU fused into the same launch
    (activated tile written back into tensor memory as the A operand of the 1x1) against the two-launch path, and
    against float64 torch for the dense case."""
    import aanet_b200.ops as ops
    B, C, H, W, deform, dil = cfg
    torch.manual_seed(29)
    x = torch.randn(B, H, W, C, device="cuda")
    idn = torch.randn(B, H, W, C, device="cuda")
    w2 = torch.randn(C, C, 3, 3, device="cuda") / (C * 9) ** 0.5
    w3 = torch.randn(C, C, 1, 1, device="cuda") / C ** 0.5
    s2, h2 = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda")
    s3, h3 = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda")
    wp2, wp3 = ops.pack_conv_weight(w2), ops.pack_conv_weight(w3)
    tail = dict(wpack=wp3, Cout=C, scale=s3, shift=h3, residual=idn, act=ops.ACT_RELU)
    if deform:
        om = torch.cat([0.7 * torch.randn(B, 36, H, W, device="cuda"),
                        2 * torch.sigmoid(torch.randn(B, 18, H, W, device="cuda"))], 1).contiguous()
        q = dict(x=x, offmask=om, om_nchw=True, wpack=wp2, Cout=C, kh=3, kw=3, scale=s2, shift=h2, act=ops.ACT_RELU,
                 stride=1, pad=dil, dil=dil, groups=1, dg=2)
    else:
        q = dict(x=x, wpack=wp2, Cout=C, kh=3, kw=3, scale=s2, shift=h2, act=ops.ACT_RELU, stride=1, pad=dil, dil=dil,
                 groups=1)
    assert ops.conv_tail_supported(dict(q, tail=tail), deform)
    fused = ops.conv_batch([dict(q, tail=tail)], deform=deform)[0]
    y2 = ops.conv_batch([q], deform=deform)[0]
    two = ops.conv2d_nhwc(y2, wp3, C, 1, 1, None, s3, h3, idn, ops.ACT_RELU)
    assert fused.shape == two.shape
    assert rel_err(npy(fused), npy(two)) < 1e-5
    if not deform:
        xr = x.permute(0, 3, 1, 2).double()
        r2 = torch.relu(torch.nn.functional.conv2d(xr, w2.double(), None, 1, dil, dil) * s2.view(1, -1, 1, 1) + h2.view(1, -1, 1, 1))
        r3 = torch.relu(torch.nn.functional.conv2d(r2, w3.double()) * s3.view(1, -1, 1, 1) + h3.view(1, -1, 1, 1)
                        + idn.permute(0, 3, 1, 2))
        assert rel_err(npy(fused.permute(0, 3, 1, 2)), npy(r3)) < 1e-5


@pytest.mark.parametrize("cfg", [
    # B, Cin, Cout, H, W  -- the CSA down-sampling convolutions (aggregation.py:353-371): 3x3, stride 2, pad 1
    (1, 64, 64, 128, 416),
    (1, 64, 32, 24, 52),
    (2, 64, 16, 17, 35),              # 16 outputs inside a 32-wide N tile, odd sizes, batch 2
    (1, 32, 16, 64, 208),
])
def test_strided_dense_tmem_vs_float64(cfg):
    """Stride-2 3x3 convolutions through the TMEM-A kernel (TMA-staged patch of every other pixel) against float64 torch
    and against the round-1 gather engine."""
    import aanet_b200.ops as ops
    B, Ci, Co, H, W = cfg
    torch.manual_seed(31)
    x = torch.randn(B, H, W, Ci, device="cuda")
    w = torch.randn(Co, Ci, 3, 3, device="cuda") / (Ci * 9) ** 0.5
    sc, sh = torch.rand(Co, device="cuda") + 0.5, torch.randn(Co, device="cuda")
    bn = max(32, ops.natural_bn(Co))
    q = dict(x=x, wpack=ops.pack_conv_weight(w, 1, bn), Cout=Co, kh=3, kw=3, scale=sc, shift=sh, act=ops.ACT_LEAKY,
             slope=0.2, stride=2, pad=1, dil=1, groups=1)
    got = ops.conv_batch([q], bn=bn)[0]
    ref = torch.nn.functional.leaky_relu(
        torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), None, 2, 1) * sc.view(1, -1, 1, 1)
        + sh.view(1, -1, 1, 1), 0.2)
    assert got.shape == (B, ref.shape[2], ref.shape[3], Co)
    assert rel_err(npy(got.permute(0, 3, 1, 2)), npy(ref)) < 1e-5
    old = os.environ.get("AANET_DENSE_TMEM")
    os.environ["AANET_DENSE_TMEM"] = "0"
    try:
        eng = ops.conv_batch([q], bn=bn)[0]
    finally:
        if old is None:
            os.environ.pop("AANET_DENSE_TMEM", None)
        else:
            os.environ["AANET_DENSE_TMEM"] = old
    assert rel_err(npy(got), npy(eng)) < 1e-5


@pytest.mark.parametrize("cfg", [
    # B, Cin, Cout, H, W, out_nchw
    (1, 64, 64, 128, 416, False),     # conv1 of the bottlenecks at the 1/3 scale
    (2, 32, 32, 33, 47, False),       # one channel block, ragged tiles, batch 2
    (1, 64, 32, 24, 52, True),        # NCHW output (non-lean epilogue)
    (1, 128, 64, 9, 20, False),       # four channel blocks
])
def test_pointwise_dense_tmem_vs_float64(cfg):
    """1x1 convolutions through the TMEM-A kernel (fewer taps than producer groups: the groups skip patch slots)
    against float64 torch and against the round-1 engine."""
    import aanet_b200.ops as ops
    B, Ci, Co, H, W, nchw = cfg
    torch.manual_seed(37)
    x = torch.randn(B, H, W, Ci, device="cuda")
    w = torch.randn(Co, Ci, 1, 1, device="cuda") / Ci ** 0.5
    sc, sh = torch.rand(Co, device="cuda") + 0.5, torch.randn(Co, device="cuda")
    wp = ops.pack_conv_weight(w)
    call = lambda: ops.conv2d_nhwc(x, wp, Co, 1, 1, None, sc, sh, None, ops.ACT_RELU, 0.0, 1, 0, 1, 1, out_nchw=nchw)
    got = call()
    ref = torch.relu(torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double()) * sc.view(1, -1, 1, 1)
                     + sh.view(1, -1, 1, 1))
    g = got if nchw else got.permute(0, 3, 1, 2)
    assert rel_err(npy(g), npy(ref)) < 1e-5
    old = os.environ.get("AANET_DENSE_TMEM")
    os.environ["AANET_DENSE_TMEM"] = "0"
    try:
        eng = call()
    finally:
        if old is None:
            os.environ.pop("AANET_DENSE_TMEM", None)
        else:
            os.environ["AANET_DENSE_TMEM"] = old
    assert rel_err(npy(got), npy(eng)) < 1e-5


@pytest.mark.parametrize("cfg", [
    # B, C, Cout, H, W, term sizes (first = output size)
    (1, 64, 64, 128, 416, [(128, 416), (64, 208), (32, 104)]),     # CSA row 0 of config 2 + conv1 of the next module
    (2, 32, 32, 33, 47, [(33, 47), (33, 47), (17, 24)]),            # row 1 pattern (two same-size terms), odd sizes
    (1, 64, 64, 24, 52, [(24, 52), (12, 26)]),
    (1, 64, 32, 20, 40, [(20, 40), (7, 13), (3, 5)]),              # odd ratios
])
def test_csa_conv1_matches_two_launches(cfg):
    """Resize-and-sum + LeakyReLU produced inside the 1x1 convolution's launch (ops.csa_conv1_nhwc) against
    csa_fuse_nhwc followed by conv2d_nhwc, and the sum against the C oracle's resize rule through csa_fuse_nhwc
    (tests/test_gpu_parity.py pins that one to the oracle)."""
    import aanet_b200.ops as ops
    B, C, Co, H, W, sizes = cfg
    torch.manual_seed(41)
    terms = [torch.randn(B, h, w, C, device="cuda") for h, w in sizes]
    w1 = torch.randn(Co, C, 1, 1, device="cuda") / C ** 0.5
    sc, sh = torch.rand(Co, device="cuda") + 0.5, torch.randn(Co, device="cuda")
    wp = ops.pack_conv_weight(w1)
    os.environ["AANET_CSA_CONV1"] = "1"          # opt-in (measured slower inside the pipeline)
    try:
        assert ops.csa_conv1_supported(terms, Co)
    finally:
        os.environ.pop("AANET_CSA_CONV1", None)
    fused, y1 = ops.csa_conv1_nhwc(terms, 0.2, wp, Co, sc, sh, ops.ACT_RELU)
    ref_sum = ops.csa_fuse_nhwc(terms, 0.2)
    ref_y1 = ops.conv2d_nhwc(ref_sum, wp, Co, 1, 1, None, sc, sh, None, ops.ACT_RELU, 0.0, 1, 0, 1, 1)
    assert fused.shape == ref_sum.shape and y1.shape == ref_y1.shape
    assert rel_err(npy(fused), npy(ref_sum)) < 1e-6
    assert rel_err(npy(y1), npy(ref_y1)) < 1e-5


@pytest.mark.parametrize("cfg", [(1, 64, 16, 64, 208), (2, 32, 16, 17, 35), (1, 64, 64, 24, 52)])
def test_strided_dense_tmem_residual_vs_float64(cfg):
    """out = LeakyReLU(bn(conv3x3 stride 2) + residual) in the TMEM-A kernel's epilogue: how the coarsest CSA row folds
    its sum into the exchange convolutions (aggregation.py:387-400)."""
    import aanet_b200.ops as ops
    B, Ci, Co, H, W = cfg
    torch.manual_seed(43)
    x = torch.randn(B, H, W, Ci, device="cuda")
    w = torch.randn(Co, Ci, 3, 3, device="cuda") / (Ci * 9) ** 0.5
    sc, sh = torch.rand(Co, device="cuda") + 0.5, torch.randn(Co, device="cuda")
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    res = torch.randn(B, Ho, Wo, Co, device="cuda")
    bn = max(32, ops.natural_bn(Co))
    q = dict(x=x, wpack=ops.pack_conv_weight(w, 1, bn), Cout=Co, kh=3, kw=3, scale=sc, shift=sh, act=ops.ACT_LEAKY,
             slope=0.2, stride=2, pad=1, dil=1, groups=1, residual=res)
    got = ops.conv_batch([q], bn=bn)[0]
    ref = torch.nn.functional.leaky_relu(
        torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), None, 2, 1) * sc.view(1, -1, 1, 1)
        + sh.view(1, -1, 1, 1) + res.permute(0, 3, 1, 2), 0.2)
    assert rel_err(npy(got.permute(0, 3, 1, 2)), npy(ref)) < 1e-5


@pytest.mark.parametrize("cfg", [(1, 64, 128, 416, [(128, 416), (64, 208), (32, 104)]), (2, 32, 33, 47, [(33, 47), (17, 24)])])
def test_csa_final_softargmin_matches_three_kernels(cfg):
    """Last aggregation module: resize-and-sum + LeakyReLU, the final 1x1 convolution (with bias) and the soft-argmin
    as ONE launch (ops.csa_conv1_nhwc, ACT_SOFTARGMIN, the sum is not written) against csa_fuse_nhwc + conv2d_nhwc +
    soft_argmin: <= 1e-4 px."""
    import aanet_b200.ops as ops
    B, C, H, W, sizes = cfg
    torch.manual_seed(47)
    terms = [torch.randn(B, h, w, C, device="cuda") for h, w in sizes]
    w1 = torch.randn(C, C, 1, 1, device="cuda") / C ** 0.5
    bias = torch.randn(C, device="cuda")
    wp = ops.pack_conv_weight(w1)
    assert ops.csa_conv1_supported(terms, C, force=True)
    none, disp = ops.csa_conv1_nhwc(terms, 0.2, wp, C, None, None, ops.ACT_SOFTARGMIN, bias=bias, keep_sum=False)
    assert none is None and disp.shape == (B, H, W)
    vol = ops.conv2d_nhwc(ops.csa_fuse_nhwc(terms, 0.2), wp, C, 1, 1, bias, None, None, None, ops.ACT_NONE, 0.0, 1, 0, 1, 1,
                          out_nchw=True)
    ref = ops.soft_argmin(vol, True)
    assert float((disp - ref).abs().max()) < 1e-4
