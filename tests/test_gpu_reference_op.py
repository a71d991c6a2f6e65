"""GPU-side second checker: the reference's OWN CUDA op (nets/deform_conv/src/*.cu|cpp, compiled by
oracle/build_ref.py into oracle/_ref/, test infrastructure only) against the sm_100a kernels, forward and
backward, at the ISA shapes.  Skipped when oracle/_ref was never built."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref_op():
    from oracle import build_ref
    try:
        mod = build_ref.load()
    except Exception as e:  # pragma: no cover
        pytest.skip("reference op not loadable: %s" % e)
    if mod is None:
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return mod


@pytest.mark.parametrize("C,H,W,B", [(64, 64, 104, 1), (32, 32, 52, 2), (16, 16, 26, 1),
                                     (64, 128, 416, 1)])        # the bench's 1/3-scale ISA layer at full size
def test_mdconv_matches_reference_cuda_op(ref_op, C, H, W, B):
    import aanet_b200.ops as ops
    torch.manual_seed(326)
    dev = "cuda"
    x = torch.randn(B, C, H, W, device=dev)
    off = 2 * torch.randn(B, 36, H, W, device=dev)
    msk = 2 * torch.sigmoid(torch.randn(B, 18, H, W, device=dev))
    w = torch.randn(C, C, 3, 3, device=dev) / (C * 9) ** 0.5
    g = torch.randn(B, C, H, W, device=dev)
    args = (3, 3, 1, 1, 2, 2, 2, 2, 1, 2, False)          # kh kw sh sw ph pw dh dw groups dg with_bias

    out_ref = torch.empty(B, C, H, W, device=dev)
    e0, e1, fake = torch.empty(0, device=dev), torch.empty(0, device=dev), torch.empty(1, device=dev)
    ref_op.modulated_deform_conv_cuda_forward(x, w, fake, e0, off, msk, out_ref, e1, *args)
    gx, goff, gm, gw, gb = (torch.zeros_like(t) for t in (x, off, msk, w, fake))
    ref_op.modulated_deform_conv_cuda_backward(x, w, fake, e0, off, msk, e1, gx, gw, gb, goff, gm, g.clone(), *args)

    xs, offs, ms, ws = (t.clone().requires_grad_() for t in (x, off, msk, w))
    out = ops.modulated_deform_conv(xs, offs, ms, ws, None, 1, 2, 2, 1, 2)
    assert rel_err(out.detach().cpu().numpy(), out_ref.cpu().numpy()) < 1e-4
    grads = torch.autograd.grad(out, (xs, offs, ms, ws), g)
    for name, a, r in zip(("gx", "goffset", "gmask", "gweight"), grads, (gx, goff, gm, gw)):
        assert rel_err(a.cpu().numpy(), r.cpu().numpy()) < 2e-4, name      # the reference sums gx with float atomics


def test_engine_mode2_full_size_matches_reference_cuda_op(ref_op):
    """The instantiation bench.py's roofline object times -- conv_umma_kernel<64, MODE 2> on channels-last input
    [1,128,416,64] with channel-plane offsets/mask (2*randn px: fractional, some outside the image), dg = 2,
    dil = 2, folded-BN affine + ReLU epilogue -- against the reference's own CUDA op
    (deform_conv_cuda_kernel.cu:570-633 + cpp:539-561) followed by the same affine + ReLU in torch."""
    import aanet_b200.ops as ops
    torch.manual_seed(326)
    dev = "cuda"
    B, C, H, W = 1, 64, 128, 416
    x = torch.randn(B, C, H, W, device=dev)
    off = 2 * torch.randn(B, 36, H, W, device=dev)
    msk = 2 * torch.sigmoid(torch.randn(B, 18, H, W, device=dev))
    w = torch.randn(C, C, 3, 3, device=dev) / (C * 9) ** 0.5
    sc, sh = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
    args = (3, 3, 1, 1, 2, 2, 2, 2, 1, 2, False)
    out_ref = torch.empty(B, C, H, W, device=dev)
    e0, e1, fake = torch.empty(0, device=dev), torch.empty(0, device=dev), torch.empty(1, device=dev)
    ref_op.modulated_deform_conv_cuda_forward(x, w, fake, e0, off, msk, out_ref, e1, *args)
    want = torch.relu(out_ref * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1))

    x_cl = x.permute(0, 2, 3, 1).contiguous()
    om = torch.cat([off, msk], 1).contiguous()                      # channel planes [B,54,H,W]
    got = ops.mdcn_nhwc(x_cl, om, ops.pack_conv_weight(w), C, 3, 3, None, sc, sh, True, 1, 2, 2, 1, 2,
                        om_nchw=True)
    assert got.shape == (B, H, W, C)
    assert rel_err(got.permute(0, 3, 1, 2).cpu().numpy(), want.cpu().numpy()) < 1e-4
    om_cl = om.permute(0, 2, 3, 1).contiguous()                     # channels-last offsets/mask
    got2 = ops.mdcn_nhwc(x_cl, om_cl, ops.pack_conv_weight(w), C, 3, 3, None, sc, sh, True, 1, 2, 2, 1, 2)
    assert rel_err(got2.permute(0, 3, 1, 2).cpu().numpy(), want.cpu().numpy()) < 1e-4
