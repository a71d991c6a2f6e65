"""GPU parity tests: the sm_100a kernels (through the C-ABI / drop-in modules) against the CPU oracle
and the golden fixtures made from the reference's own code.

Tolerances (BASELINE.json north_star): fp32 volumes within 1e-4 relative (max|a-b| / max|b|),
disparity within 1e-3 px max abs.  Gradients: 1e-4 relative against the float64-accumulating oracle
(the reference's own grad_input uses float atomics, SURVEY.md section 7 "Backward determinism").
"""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

VOL_TOL = 1e-4
GRAD_TOL = 1e-4
DISP_TOL = 1e-3


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def npy(t):
    return t.detach().float().cpu().numpy()


@pytest.fixture(scope="module")
def ops():
    import aanet_b200.ops as o
    return o


# ------------------------------------------------------------------------------------ correlation
@pytest.mark.parametrize("tag", ["a", "narrow", "c1", "c128"])
def test_corr_golden(ops, golden, tag):
    z = golden("corr")
    L, R = cu(z[tag + "_L"]).requires_grad_(), cu(z[tag + "_R"]).requires_grad_()
    out = ops.correlation(L, R, int(z[tag + "_D"]))
    assert rel_err(npy(out), z[tag + "_out"]) < VOL_TOL
    gL, gR = torch.autograd.grad(out, (L, R), cu(z[tag + "_g"]))
    assert rel_err(npy(gL), z[tag + "_gL"]) < GRAD_TOL
    assert rel_err(npy(gR), z[tag + "_gR"]) < GRAD_TOL


@pytest.mark.parametrize("shape", [(1, 32, 6, 200, 64), (2, 16, 3, 131, 96), (1, 8, 4, 30, 48),
                                   (1, 128, 2, 416, 64), (3, 5, 1, 257, 7),
                                   (1, 128, 3, 208, 32), (1, 128, 2, 104, 16), (1, 64, 2, 300, 128),   # tcgen05 widths
                                   (1, 33, 2, 140, 20), (2, 8, 2, 100, 100),
                                   (1, 16, 2, 300, 144), (1, 6, 3, 203, 130)])                        # D > 128: FFMA kernels
def test_corr_oracle(ops, shape):
    B, C, H, W, D = shape
    rng = np.random.default_rng(326)
    L = np.maximum(rng.standard_normal((B, C, H, W)), 0).astype(np.float32)
    R = np.maximum(rng.standard_normal((B, C, H, W)), 0).astype(np.float32)
    g = rng.standard_normal((B, D, H, W)).astype(np.float32)
    Lc, Rc = cu(L).requires_grad_(), cu(R).requires_grad_()
    out = ops.correlation(Lc, Rc, D)
    ref = orc.corr_fwd(L, R, D)
    assert rel_err(npy(out), ref) < VOL_TOL
    o = npy(out)
    for d in range(1, D):      # the w<d triangle is exactly zero (cost.py:41 new_zeros)
        assert np.all(o[:, d, :, :min(d, W)] == 0.0)
    gL, gR = torch.autograd.grad(out, (Lc, Rc), cu(g))
    rL, rR = orc.corr_bwd(L, R, g)
    assert rel_err(npy(gL), rL) < GRAD_TOL
    assert rel_err(npy(gR), rR) < GRAD_TOL


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_corr_bwd_second_device_large_smem(ops):
    """One process driving two GPUs (the reference's nn.DataParallel scripts): the backward kernel's > 48 KB
    dynamic shared memory opt-in is per device."""
    rng = np.random.default_rng(5)
    L = np.maximum(rng.standard_normal((1, 8, 2, 150)), 0).astype(np.float32)
    R = np.maximum(rng.standard_normal((1, 8, 2, 150)), 0).astype(np.float32)
    g = rng.standard_normal((1, 96, 2, 150)).astype(np.float32)
    rL, rR = orc.corr_bwd(L, R, g)
    for dev in ("cuda:0", "cuda:1"):
        Lc = torch.from_numpy(L).to(dev).requires_grad_(); Rc = torch.from_numpy(R).to(dev).requires_grad_()
        out = ops.correlation(Lc, Rc, 96)
        gL, gR = torch.autograd.grad(out, (Lc, Rc), torch.from_numpy(g).to(dev))
        assert rel_err(npy(gL), rL) < GRAD_TOL and rel_err(npy(gR), rR) < GRAD_TOL


def test_corr_full_size_properties(ops):
    """KITTI 1/3 scale (BASELINE config 2): linearity in L, zero band, d=0 plane = channel mean."""
    torch.manual_seed(326)
    L = torch.relu(torch.randn(1, 128, 128, 416, device="cuda"))
    L2 = torch.relu(torch.randn(1, 128, 128, 416, device="cuda"))
    R = torch.relu(torch.randn(1, 128, 128, 416, device="cuda"))
    a, b2 = ops.correlation(L, R, 64), ops.correlation(L2, R, 64)
    s = ops.correlation(L + 2 * L2, R, 64)
    assert rel_err(npy(s), npy(a + 2 * b2)) < 1e-5
    assert rel_err(npy(a[:, 0]), npy((L * R).mean(1))) < 1e-5
    for d in (1, 17, 63):
        assert torch.all(a[:, d, :, :d] == 0)
        assert rel_err(npy(a[:, d, :, d:]), npy((L[..., d:] * R[..., :-d]).mean(1))) < 1e-5


def test_corr_bf16_variant(ops):
    """Config-5 variant: bf16 features.  Exact w.r.t. the oracle fed the SAME bf16-rounded features (fp32
    accumulation), and within a stated error of the fp32 volume: rel 1e-2 (8 mantissa bits per factor)."""
    torch.manual_seed(326)
    B, C, H, W, D = 1, 128, 16, 416, 64
    L = torch.relu(torch.randn(B, C, H, W, device="cuda"))
    R = torch.relu(torch.randn(B, C, H, W, device="cuda"))
    Lb, Rb = L.bfloat16(), R.bfloat16()
    vol = ops.correlation(Lb, Rb, D)
    assert vol.dtype == torch.float32
    ref_same = orc.corr_fwd(npy(Lb), npy(Rb), D)
    assert rel_err(npy(vol), ref_same) < 1e-5
    assert rel_err(npy(vol), npy(ops.correlation(L, R, D))) < 1e-2
    assert torch.all(vol[:, 63, :, :63] == 0)
    with pytest.raises(RuntimeError):
        ops.correlation(Lb[..., :412].contiguous(), Rb[..., :412].contiguous(), D)     # W % 8 != 0


# ------------------------------------------------------------------------------------ soft-argmin
@pytest.mark.parametrize("tag", ["sim", "cost", "d1", "peaky"])
def test_softargmin_golden(ops, golden, tag):
    z = golden("softargmin")
    c = cu(z[tag + "_cost"]).requires_grad_()
    disp = ops.soft_argmin(c, bool(z[tag + "_sim"]))
    assert np.abs(npy(disp) - z[tag + "_disp"]).max() < DISP_TOL
    gc, = torch.autograd.grad(disp, c, cu(z[tag + "_g"]))
    assert rel_err(npy(gc), z[tag + "_gcost"]) < GRAD_TOL


@pytest.mark.parametrize("shape", [(1, 64, 128, 416), (2, 32, 7, 9), (1, 16, 32, 104), (1, 96, 5, 13),
                                   (1, 3, 1, 1)])
@pytest.mark.parametrize("sim", [True, False])
def test_softargmin_oracle(ops, shape, sim):
    rng = np.random.default_rng(7)
    c = (rng.standard_normal(shape) * 4).astype(np.float32)
    g = rng.standard_normal((shape[0],) + shape[2:]).astype(np.float32)
    cc = cu(c).requires_grad_()
    disp = ops.soft_argmin(cc, sim)
    assert np.abs(npy(disp) - orc.softargmin_fwd(c, sim)).max() < DISP_TOL
    gc, = torch.autograd.grad(disp, cc, cu(g))
    assert rel_err(npy(gc), orc.softargmin_bwd(c, g, sim)) < GRAD_TOL


def test_softargmin_properties(ops):
    torch.manual_seed(1)
    c = torch.randn(1, 64, 128, 416, device="cuda") * 3
    d0 = ops.soft_argmin(c, True)
    assert torch.all((d0 >= 0) & (d0 <= 63))
    assert (ops.soft_argmin(c + 5.0, True) - d0).abs().max() < 1e-3      # shift invariance
    onehot = torch.full((1, 64, 4, 8), -1e4, device="cuda")
    onehot[:, 37] = 0
    assert (ops.soft_argmin(onehot, True) - 37).abs().max() < 1e-5


# ------------------------------------------------------------------------------------ mdconv
def _mdcn_case(z, tag):
    st, pad, dil, grp, dg, has_b = [int(v) for v in z[tag + "_cfg"]]
    return dict(stride=st, pad=pad, dil=dil, groups=grp, dg=dg, has_b=has_b)


@pytest.fixture(params=["umma", "generic"])
def mdcn_path(request, ops):
    """Run every mdconv forward test through both the tcgen05 engine (when the shape qualifies) and the
    shape-generic FFMA kernel (workspace = NULL)."""
    ops.FORCE_GENERIC_MDCN = request.param == "generic"
    yield request.param
    ops.FORCE_GENERIC_MDCN = False


@pytest.mark.parametrize("tag", ["isa", "s2", "grp", "far", "k1", "v1"])
def test_mdcn_golden(ops, golden, tag, mdcn_path):
    z = golden("mdcn")
    c = _mdcn_case(z, tag)
    x = cu(z[tag + "_x_f64"]).requires_grad_()
    off = cu(z[tag + "_off_f64"]).requires_grad_()
    w = cu(z[tag + "_w_f64"]).requires_grad_()
    b = cu(z[tag + "_b_f64"]).requires_grad_() if c["has_b"] else None
    if tag == "v1":
        out = ops.deform_conv(x, off, w, c["stride"], c["pad"], c["dil"], c["groups"], c["dg"])
        params = (x, off, w)
    else:
        m = cu(z[tag + "_mask_f64"]).requires_grad_()
        out = ops.modulated_deform_conv(x, off, m, w, b, c["stride"], c["pad"], c["dil"], c["groups"], c["dg"])
        params = (x, off, m, w) + ((b,) if b is not None else ())
    assert rel_err(npy(out), z[tag + "_out_f64"]) < VOL_TOL
    grads = torch.autograd.grad(out, params, cu(z[tag + "_g_f64"]))
    names = ["gx", "goff", "gw"] if tag == "v1" else ["gx", "goff", "gmask", "gw"] + (["gb"] if b is not None else [])
    for n, g in zip(names, grads):
        assert rel_err(npy(g), z["%s_%s_f64" % (tag, n)]) < GRAD_TOL, n


@pytest.mark.parametrize("cfg", [
    # B, Cin, Cout, H, W, stride, dil, groups, dg, bias
    (1, 64, 64, 24, 52, 1, 2, 1, 2, False),      # ISA shape, small plane
    (2, 32, 32, 16, 26, 1, 2, 1, 2, False),
    (1, 16, 16, 8, 13, 1, 2, 1, 2, False),
    (1, 24, 40, 15, 17, 2, 1, 1, 4, True),       # stride 2, Cin != Cout
    (1, 96, 96, 9, 11, 1, 2, 1, 8, False),
    (2, 12, 18, 10, 10, 1, 1, 3, 2, True),       # conv groups straddling deformable groups
    (1, 128, 128, 6, 20, 1, 2, 1, 1, False),     # two output tiles
    (1, 24, 24, 7, 9, 1, 2, 1, 2, False),        # tcgen05 backward with B*K*dg*Ho*Wo % 4 != 0 (partial warps in the scatter)
])
def test_mdcn_oracle(ops, cfg, mdcn_path):
    B, Ci, Co, H, W, st, dil, grp, dg, bias = cfg
    rng = np.random.default_rng(11)
    k, pad = 3, dil
    Ho, Wo = orc.mdcn_out_hw(H, W, k, st, pad, dil)
    x = rng.standard_normal((B, Ci, H, W)).astype(np.float32)
    off = (2 * rng.standard_normal((B, dg * 18, Ho, Wo))).astype(np.float32)
    msk = (2 / (1 + np.exp(-rng.standard_normal((B, dg * 9, Ho, Wo))))).astype(np.float32)
    w = (rng.standard_normal((Co, Ci // grp, k, k)) / np.sqrt(Ci * 9)).astype(np.float32)
    b = rng.standard_normal(Co).astype(np.float32) if bias else None
    g = rng.standard_normal((B, Co, Ho, Wo)).astype(np.float32)
    xc, oc, mc, wc = [cu(a).requires_grad_() for a in (x, off, msk, w)]
    bc = cu(b).requires_grad_() if bias else None
    out = ops.modulated_deform_conv(xc, oc, mc, wc, bc, st, pad, dil, grp, dg)
    ref = orc.mdcn_fwd(x, off, msk, w, b, st, pad, dil, grp, dg)
    assert rel_err(npy(out), ref) < VOL_TOL
    grads = torch.autograd.grad(out, (xc, oc, mc, wc) + ((bc,) if bias else ()), cu(g))
    rg = orc.mdcn_bwd(x, off, msk, w, g, bias, st, pad, dil, grp, dg)
    for n, a, r in zip(["gx", "goff", "gmask", "gw", "gb"], grads, rg):
        assert rel_err(npy(a), r) < GRAD_TOL, n


def test_mdcn_zero_offset_is_dilated_conv(ops, mdcn_path):
    """Full ISA size: with offset = 0 and mask = 1 the op is an ordinary dilated conv (deform.py:75-76)."""
    torch.manual_seed(3)
    x = torch.randn(1, 64, 128, 416, device="cuda")
    w = torch.randn(64, 64, 3, 3, device="cuda") / 24
    off = torch.zeros(1, 36, 128, 416, device="cuda")
    m = torch.ones(1, 18, 128, 416, device="cuda")
    out = ops.modulated_deform_conv(x, off, m, w, None, 1, 2, 2, 1, 2)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = torch.nn.functional.conv2d(x, w, padding=2, dilation=2)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert rel_err(npy(out), npy(ref)) < VOL_TOL


def test_mdcn_fused_epilogue(ops):
    torch.manual_seed(5)
    x = torch.randn(2, 16, 9, 12, device="cuda")
    off = torch.randn(2, 36, 9, 12, device="cuda")
    m = torch.rand(2, 18, 9, 12, device="cuda") * 2
    w = torch.randn(16, 16, 3, 3, device="cuda") / 12
    sc, sh = torch.rand(16, device="cuda") + 0.5, torch.randn(16, device="cuda")
    plain = ops.modulated_deform_conv(x, off, m, w, None, 1, 2, 2, 1, 2)
    fused = ops.modulated_deform_conv_fused(x, off, m, w, None, 1, 2, 2, 1, 2, sc, sh, True)
    ref = torch.relu(plain * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1))
    assert rel_err(npy(fused), npy(ref)) < 1e-5


# ------------------------------------------------------------------------------------ dense conv engine
@pytest.mark.parametrize("cfg", [
    # B, Cin, Cout, H, W, k, stride, pad, dil, groups
    (1, 64, 64, 128, 416, 1, 1, 0, 1, 1),     # ISA conv1 / conv3 at the 1/3 scale
    (1, 64, 64, 40, 52, 3, 1, 1, 1, 1),       # SimpleBottleneck conv2
    (2, 64, 54, 24, 40, 3, 1, 2, 2, 2),       # offset_conv: grouped, dilated, bias
    (1, 64, 32, 33, 47, 3, 2, 1, 1, 1),       # CSA down path, stride 2, odd size
    (1, 16, 64, 16, 26, 1, 1, 0, 1, 1),       # CSA up path 1x1
    (2, 32, 32, 9, 300, 3, 1, 1, 1, 1),
    (1, 8, 24, 7, 9, 3, 1, 1, 1, 1),
    (1, 128, 128, 12, 20, 3, 1, 1, 1, 1),     # BN = 128
    (1, 96, 96, 10, 14, 1, 1, 0, 1, 1),       # BN = 96
])
def test_conv2d_fused(ops, cfg):
    B, Ci, Co, H, W, k, st, pad, dil, grp = cfg
    torch.manual_seed(13)
    x = torch.randn(B, Ci, H, W, device="cuda")
    w = torch.randn(Co, Ci // grp, k, k, device="cuda") / (Ci * k * k / grp) ** 0.5
    bias = torch.randn(Co, device="cuda")
    scale, shift = torch.rand(Co, device="cuda") + 0.5, torch.randn(Co, device="cuda")
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = torch.nn.functional.conv2d(x.double(), w.double(), bias.double(), st, pad, dil, grp)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    res = torch.randn_like(ref).float()
    out = ops.conv2d_fused(x, w, bias, None, None, None, ops.ACT_NONE, 0.2, st, pad, dil, grp)
    assert rel_err(npy(out), npy(ref)) < 1e-5
    full = ops.conv2d_fused(x, w, bias, scale, shift, res, ops.ACT_LEAKY, 0.2, st, pad, dil, grp)
    ref2 = torch.nn.functional.leaky_relu(ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1) + res, 0.2)
    assert rel_err(npy(full), npy(ref2)) < 1e-5
    relu = ops.conv2d_fused(x, w, None, scale, shift, None, ops.ACT_RELU, 0.0, st, pad, dil, grp)
    ref3 = torch.relu((ref - bias.view(1, -1, 1, 1).double()) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    assert rel_err(npy(relu), npy(ref3)) < 1e-5


# ------------------------------------------------------------------------------------ channels-last engine calls
def test_layout_transposes(ops):
    x = torch.randn(2, 20, 7, 45, device="cuda")
    t = ops.nchw_to_nhwc(x)
    assert t.shape == (2, 7, 45, 20) and torch.equal(t, x.permute(0, 2, 3, 1).contiguous())
    assert torch.equal(ops.nhwc_to_nchw(t), x)


@pytest.mark.parametrize("cfg", [
    (1, 64, 64, 128, 416, 1, 1, 0, 1, 1), (2, 64, 54, 24, 40, 3, 1, 2, 2, 2), (1, 64, 32, 33, 47, 3, 2, 1, 1, 1),
    (3, 16, 16, 19, 23, 3, 1, 1, 1, 1), (1, 32, 64, 64, 208, 1, 1, 0, 1, 1), (1, 8, 4, 5, 300, 3, 1, 1, 1, 1),
])
def test_conv2d_nhwc(ops, cfg):
    B, Ci, Co, H, W, k, st, pad, dil, grp = cfg
    torch.manual_seed(17)
    x = torch.randn(B, Ci, H, W, device="cuda")
    w = torch.randn(Co, Ci // grp, k, k, device="cuda") / (Ci * k * k / grp) ** 0.5
    bias, scale, shift = torch.randn(Co, device="cuda"), torch.rand(Co, device="cuda") + 0.5, torch.randn(Co, device="cuda")
    ref = torch.nn.functional.conv2d(x.double(), w.double(), bias.double(), st, pad, dil, grp)
    res = torch.randn_like(ref).float()
    wp = ops.pack_conv_weight(w, grp)
    xt = ops.nchw_to_nhwc(x)
    out = ops.conv2d_nhwc(xt, wp, Co, k, k, bias, scale, shift, ops.nchw_to_nhwc(res), ops.ACT_LEAKY, 0.2, st, pad, dil, grp)
    ref2 = torch.nn.functional.leaky_relu(ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1) + res, 0.2)
    assert rel_err(npy(out.permute(0, 3, 1, 2)), npy(ref2)) < 1e-5
    out_nchw = ops.conv2d_nhwc(xt, wp, Co, k, k, bias, None, None, res, ops.ACT_RELU, 0.0, st, pad, dil, grp, out_nchw=True)
    assert rel_err(npy(out_nchw), npy(torch.relu(ref + res))) < 1e-5
    # offset/mask head: channels >= n_off get 2*sigmoid
    n_off = (Co * 2) // 3
    om = ops.conv2d_nhwc(xt, wp, Co, k, k, bias, None, None, None, ops.ACT_OFFSET_MASK, 0.0, st, pad, dil, grp,
                         n_offset_ch=n_off, mask_scale=2.0)
    ref_om = ref.clone()
    ref_om[:, n_off:] = 2 * torch.sigmoid(ref[:, n_off:])
    assert rel_err(npy(om.permute(0, 3, 1, 2)), npy(ref_om)) < 1e-5


@pytest.mark.parametrize("cfg", [(1, 64, 64, 24, 52, 1, 2, 2), (2, 32, 32, 16, 26, 1, 2, 2), (1, 16, 24, 9, 13, 2, 1, 2),
                                 (1, 96, 96, 9, 11, 1, 2, 8), (1, 8, 8, 12, 10, 1, 2, 2)])
def test_mdcn_nhwc(ops, cfg):
    B, Ci, Co, H, W, st, dil, dg = cfg
    rng = np.random.default_rng(23)
    Ho, Wo = orc.mdcn_out_hw(H, W, 3, st, dil, dil)
    x = rng.standard_normal((B, Ci, H, W)).astype(np.float32)
    off = (2 * rng.standard_normal((B, dg * 18, Ho, Wo))).astype(np.float32)
    msk = rng.uniform(0, 2, (B, dg * 9, Ho, Wo)).astype(np.float32)
    w = (rng.standard_normal((Co, Ci, 3, 3)) / np.sqrt(Ci * 9)).astype(np.float32)
    ref = orc.mdcn_fwd(x, off, msk, w, None, st, dil, dil, 1, dg)
    om = ops.nchw_to_nhwc(cu(np.concatenate([off, msk], 1)))
    out = ops.mdcn_nhwc(ops.nchw_to_nhwc(cu(x)), om, ops.pack_conv_weight(cu(w)), Co, 3, 3, None, None, None, False,
                        st, dil, dil, 1, dg)
    assert rel_err(npy(out.permute(0, 3, 1, 2)), ref) < VOL_TOL
    planes = ops.mdcn_nhwc(ops.nchw_to_nhwc(cu(x)), cu(np.concatenate([off, msk], 1)), ops.pack_conv_weight(cu(w)),
                           Co, 3, 3, None, None, None, False, st, dil, dil, 1, dg, om_nchw=True)
    assert rel_err(npy(planes.permute(0, 3, 1, 2)), ref) < VOL_TOL
    v1 = ops.mdcn_nhwc(ops.nchw_to_nhwc(cu(x)), ops.nchw_to_nhwc(cu(off)), ops.pack_conv_weight(cu(w)), Co, 3, 3,
                       None, None, None, False, st, dil, dil, 1, dg, out_nchw=True)          # DCNv1: no mask
    assert rel_err(npy(v1), orc.mdcn_fwd(x, off, None, w, None, st, dil, dil, 1, dg)) < VOL_TOL


def test_conv_batch_three_problems(ops):
    """One persistent launch walking the tiles of three different dense problems (1x1 / 3x3 stride 2 / 3x3,
    different channel counts and sizes, shared N-tile width 64), and of three deformable problems."""
    torch.manual_seed(29)
    specs = [(64, 64, 40, 56, 1, 1, 0, 1), (64, 32, 33, 47, 3, 2, 1, 1), (16, 16, 20, 28, 3, 1, 1, 1)]
    probs, refs = [], []
    for Ci, Co, H, W, k, st, pad, dil in specs:
        x = torch.randn(2, Ci, H, W, device="cuda")
        w = torch.randn(Co, Ci, k, k, device="cuda") / (Ci * k * k) ** 0.5
        b = torch.randn(Co, device="cuda")
        refs.append(torch.relu(torch.nn.functional.conv2d(x.double(), w.double(), b.double(), st, pad, dil)))
        probs.append(dict(x=ops.nchw_to_nhwc(x), wpack=ops.pack_conv_weight(w, 1, 64), Cout=Co, kh=k, kw=k, bias=b,
                          act=ops.ACT_RELU, stride=st, pad=pad, dil=dil))
    outs = ops.conv_batch(probs, deform=False, bn=64)
    for o, r in zip(outs, refs):
        assert rel_err(npy(o.permute(0, 3, 1, 2)), npy(r)) < 1e-5
    rng = np.random.default_rng(31)
    dprobs, drefs = [], []
    for C, H, W in [(64, 24, 40), (32, 12, 20), (16, 6, 10)]:
        x = rng.standard_normal((1, C, H, W)).astype(np.float32)
        off = (2 * rng.standard_normal((1, 36, H, W))).astype(np.float32)
        msk = rng.uniform(0, 2, (1, 18, H, W)).astype(np.float32)
        w = (rng.standard_normal((C, C, 3, 3)) / np.sqrt(C * 9)).astype(np.float32)
        drefs.append(orc.mdcn_fwd(x, off, msk, w, None, 1, 2, 2, 1, 2))
        dprobs.append(dict(x=ops.nchw_to_nhwc(cu(x)), offmask=ops.nchw_to_nhwc(cu(np.concatenate([off, msk], 1))),
                           wpack=ops.pack_conv_weight(cu(w), 1, 64), Cout=C, kh=3, kw=3, stride=1, pad=2, dil=2, dg=2))
    for o, r in zip(ops.conv_batch(dprobs, deform=True, bn=64), drefs):
        assert rel_err(npy(o.permute(0, 3, 1, 2)), r) < VOL_TOL


@pytest.mark.parametrize("cfg", [
    ((2, 8, 31, 45), [(31, 45), (16, 23), (8, 12)]),          # odd ratios, partial tiles
    ((1, 64, 128, 416), [(128, 416), (64, 208), (32, 104)]),  # out0 of config 2: exact 2x / 4x
    ((1, 32, 64, 208), [(64, 208), (64, 208), (32, 104)]),    # out1
    ((1, 16, 32, 104), [(32, 104), (32, 104), (32, 104)]),    # out2: same-size terms only
    ((2, 12, 9, 70), [(9, 70), (3, 11)]),                     # C/4 not a power of two, two terms
    ((1, 4, 5, 7), [(5, 7), (1, 1)]),                         # single-pixel source
    ((1, 8, 6, 10), [(6, 10), (12, 20)]),                     # downsampling term -> per-pixel kernel
    ((1, 640, 8, 40), [(8, 40), (4, 20), (2, 10)]),           # patches exceed the smem budget -> per-pixel kernel
])
def test_csa_fuse_nhwc(ops, cfg):
    rng = np.random.default_rng(4)
    (B, C, H, W), ths = cfg
    terms = [rng.standard_normal((B, C, h, w)).astype(np.float32) for h, w in ths]
    out = ops.csa_fuse_nhwc([ops.nchw_to_nhwc(cu(t)) for t in terms], 0.2)
    assert rel_err(npy(out.permute(0, 3, 1, 2)), orc.csa_fuse_fwd(terms, (H, W), 0.2)) < 1e-5


# ------------------------------------------------------------------------------------ CSA fuse
@pytest.mark.parametrize("tag", ["x2x4", "odd", "same", "two"])
def test_csa_golden(ops, golden, tag):
    z = golden("csa")
    n = int(z[tag + "_n"])
    terms = [cu(z["%s_t%d" % (tag, i)]).requires_grad_() for i in range(n)]
    out = ops.csa_fuse(terms, 0.2)
    assert rel_err(npy(out), z[tag + "_out"]) < 1e-5
    grads = torch.autograd.grad(out, terms, cu(z[tag + "_g"]))
    for i in range(n):
        assert rel_err(npy(grads[i]), z["%s_gt%d" % (tag, i)]) < 1e-5


@pytest.mark.parametrize("shape", [((1, 64, 128, 416), [(128, 416), (64, 208), (32, 104)]),
                                   ((2, 5, 31, 45), [(31, 45), (16, 23), (8, 12)]),
                                   ((1, 16, 32, 104), [(32, 104), (32, 104), (32, 104)])])
def test_csa_oracle(ops, shape):
    (B, C, H, W), ths = shape
    rng = np.random.default_rng(2)
    terms = [rng.standard_normal((B, C, h, w)).astype(np.float32) for h, w in ths]
    g = rng.standard_normal((B, C, H, W)).astype(np.float32)
    tc = [cu(t).requires_grad_() for t in terms]
    out = ops.csa_fuse(tc, 0.2)
    ref = orc.csa_fuse_fwd(terms, (H, W), 0.2)
    assert rel_err(npy(out), ref) < 1e-5
    grads = torch.autograd.grad(out, tc, cu(g))
    rg = orc.csa_fuse_bwd(ref, g, ths, 0.2)
    for a, r in zip(grads, rg):
        assert rel_err(npy(a), r) < 1e-5


# ------------------------------------------------------------------------------------ error behaviour
def test_errors(ops):
    x = torch.randn(1, 4, 5, 5)
    with pytest.raises(NotImplementedError):          # deform_conv.py:135-136
        ops.modulated_deform_conv(x, x, x, x, None, 1, 1, 1, 1, 1)
    xc = torch.randn(1, 4, 5, 5, device="cuda")
    with pytest.raises(RuntimeError):                 # cpp:512-516 channel mismatch
        ops.modulated_deform_conv(xc, torch.zeros(1, 18, 5, 5, device="cuda"),
                                  torch.ones(1, 9, 5, 5, device="cuda"),
                                  torch.randn(4, 3, 3, 3, device="cuda"), None, 1, 1, 1, 1, 1)
    with pytest.raises(TypeError):
        ops.soft_argmin(xc.double(), True)


# ------------------------------------------------------------------------------------ refinement front end
@pytest.mark.parametrize("tag", ["x3", "same", "odd", "x2"])
def test_refine_frontend_golden(ops, golden, tag):
    """Fused upsample + rescale + disp_warp + error + concat vs tensors captured inside the reference's
    StereoDRNetRefinement.forward (refinement.py:80-95, warp.py:41-64)."""
    z = golden("refinement")
    concat, disp = ops.refine_frontend(cu(z[tag + "_low"]), cu(z[tag + "_left"]), cu(z[tag + "_right"]))
    assert np.abs(npy(disp) - z[tag + "_disp"]).max() < 1e-5 * max(1.0, np.abs(z[tag + "_disp"]).max())
    assert np.abs(npy(concat) - z[tag + "_concat"]).max() < 1e-5
    assert np.array_equal(npy(concat)[:, 3:], z[tag + "_left"])


def test_refine_frontend_full_size_vs_oracle_and_torch(ops):
    """KITTI size (low_disp at 1/3 -> 384x1248): against the oracle, and against the differentiable torch
    composite the modules use when autograd is on (same arithmetic as the reference, minus its sync)."""
    from aanet_b200.nets.refine import refine_frontend_torch
    torch.manual_seed(326)
    low = torch.rand(1, 128, 416, device="cuda") * 60
    left, right = torch.rand(1, 3, 384, 1248, device="cuda"), torch.rand(1, 3, 384, 1248, device="cuda")
    concat, disp = ops.refine_frontend(low, left, right)
    # same device, same arithmetic as the reference (F.interpolate + grid_sample): disparity within 1e-3 px
    tc, td = refine_frontend_torch(low, left, right)
    assert (td.detach() - disp).abs().max() < 1e-3 and (tc.detach() - concat).abs().max() < 1e-3
    # CPU oracle: the upsampling source index is an FMA on the GPU (as in ATen's CUDA kernel) and two roundings on
    # the CPU, which moves a ~180 px disparity by up to a few 1e-3 px: relative bound
    rc, rd = orc.refine_frontend_fwd(npy(low), npy(left), npy(right))
    assert np.abs(npy(disp) - rd).max() < 1e-4 * np.abs(rd).max()
    assert np.abs(npy(concat) - rc).max() < 1e-2 and np.abs(npy(concat) - rc).mean() < 1e-4
    with pytest.raises(RuntimeError):
        ops.refine_frontend(low[:, :100], left[..., :416].contiguous(), right[..., :416].contiguous())   # W == w, H != h


@pytest.mark.parametrize("shape", [(1, 128, 3, 416, 64), (2, 32, 2, 208, 32), (1, 16, 2, 104, 16), (1, 20, 3, 131, 96),
                                   (1, 8, 2, 300, 128), (1, 12, 2, 70, 4)])
def test_corr_nhwc_equals_nchw(ops, shape):
    """Channels-last volume of the fused path == the reference-layout volume, element for element."""
    B, C, H, W, D = shape
    torch.manual_seed(5)
    L = torch.relu(torch.randn(B, C, H, W, device="cuda"))
    R = torch.relu(torch.randn(B, C, H, W, device="cuda"))
    a = ops.correlation(L, R, D)
    b = ops.correlation_nhwc(L, R, D)
    assert b.shape == (B, H, W, D) and torch.equal(b.permute(0, 3, 1, 2), a)
    with pytest.raises(RuntimeError):
        ops.correlation_nhwc(L, R, 130)
    with pytest.raises(RuntimeError):
        ops.correlation_nhwc(L, R, 6)


# ------------------------------------------------------------------------------------ 5-D volumes (cost.py:22-38)
@pytest.mark.parametrize("kind", ["difference", "concat"])
@pytest.mark.parametrize("tag", ["a", "narrow"])
def test_cost5d_golden(golden, kind, tag):
    import aanet_b200.nets as n
    z = golden("cost5d")
    k = "%s_%s_" % (kind, tag)
    L, R = cu(z[k + "L"]).requires_grad_(), cu(z[k + "R"]).requires_grad_()
    out = n.CostVolume(int(z[k + "D"]), kind)(L, R)
    assert np.array_equal(npy(out), z[k + "out"])                 # a subtraction / a copy: bit-exact
    gL, gR = torch.autograd.grad(out, (L, R), cu(z[k + "g"]))
    assert rel_err(npy(gL), z[k + "gL"]) < 1e-5 and rel_err(npy(gR), z[k + "gR"]) < 1e-5


@pytest.mark.parametrize("kind", ["difference", "concat"])
@pytest.mark.parametrize("shape", [(1, 32, 20, 52, 48), (2, 5, 3, 31, 7)])
def test_cost5d_oracle(kind, shape):
    """StereoNet/PSMNet-like sizes (W % 4 == 0: 128-bit stores) and ragged ones, plus the pyramid wrapper."""
    import aanet_b200.nets as n
    B, C, H, W, D = shape
    rng = np.random.default_rng(7)
    L = rng.standard_normal((B, C, H, W)).astype(np.float32)
    R = rng.standard_normal((B, C, H, W)).astype(np.float32)
    Lc, Rc = cu(L).requires_grad_(), cu(R).requires_grad_()
    out = n.CostVolume(D, kind)(Lc, Rc)
    assert np.array_equal(npy(out), orc.cost5d_fwd(L, R, D, kind))
    g = rng.standard_normal(out.shape).astype(np.float32)
    gL, gR = torch.autograd.grad(out, (Lc, Rc), cu(g))
    rL, rR = orc.cost5d_bwd(g, kind)
    assert rel_err(npy(gL), rL) < 1e-5 and rel_err(npy(gR), rR) < 1e-5
    pyr = n.CostVolumePyramid(D, kind)([cu(L), cu(L[..., ::2, ::2])], [cu(R), cu(R[..., ::2, ::2])])
    assert np.array_equal(npy(pyr[1]), orc.cost5d_fwd(L[..., ::2, ::2], R[..., ::2, ::2], D // 2, kind))
