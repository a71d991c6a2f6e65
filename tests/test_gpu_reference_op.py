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


@pytest.mark.parametrize("C,H,W,B", [(64, 64, 104, 1), (32, 32, 52, 2), (16, 16, 26, 1)])
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
