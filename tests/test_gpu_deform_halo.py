"""Deformable ISA kernel with the gather served from a TMA-staged halo (aanet_b200/csrc/deform_halo.cu) against
the C oracle (reference deform_conv_cuda_kernel.cu:570-633 restated) and against the round-1 gather engine,
including offsets far larger than the staged margin (global fallback) and every margin setting."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


VARIANTS = {"smem-halo": "AANET_DEFORM_HALO", "tmem": "AANET_DEFORM_TMEM"}
KEYS = ("AANET_DEFORM_HALO", "AANET_DEFORM_TMEM", "AANET_DEFORM_MARGIN", "AANET_DEFORM_ROWS")


def _select(variant):
    for k in VARIANTS.values():
        os.environ[k] = "0"
    if variant is not None:
        os.environ[VARIANTS[variant]] = "1"


@pytest.fixture(params=["smem-halo", "tmem"])
def deform_halo(request):
    """Selects one of the two halo-staged deformable kernels: A operand through shared memory (deform_halo.cu) or
    written straight into tensor memory (deform_tmem.cu)."""
    old = {k: os.environ.get(k) for k in KEYS}
    _select(request.param)
    yield request.param
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def _run(ops, x, off, msk, w, sc, sh, dg, dil, om_nchw=True):
    B, C, H, W = x.shape
    Co = w.shape[0]
    x_cl = x.permute(0, 2, 3, 1).contiguous()
    om = torch.cat([off, msk], 1).contiguous()
    if not om_nchw:
        om = om.permute(0, 2, 3, 1).contiguous()
    out = ops.mdcn_nhwc(x_cl, om, ops.pack_conv_weight(w), Co, 3, 3, None, sc, sh, True, 1, dil, dil, 1, dg,
                        om_nchw=om_nchw)
    return out.permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("cfg", [
    # B, Cin, Cout, H, W, dg, dil, offset sigma, margin
    (1, 64, 64, 24, 52, 2, 2, 2.0, None),      # ISA shape, small plane, default margin
    (2, 64, 64, 17, 35, 2, 2, 1.0, 2),         # ragged tiles, batch 2
    (1, 64, 64, 20, 40, 2, 2, 8.0, 2),         # most samples leave the staged patch: global fallback
    (1, 64, 64, 20, 40, 2, 2, 2.0, 0),         # margin 0: only the regular grid is staged
    (1, 64, 64, 20, 40, 2, 2, 2.0, 1),
    (1, 128, 128, 12, 20, 2, 2, 2.0, None),    # Cd = 64 (two channel blocks per deformable group), two N tiles
    (1, 128, 64, 9, 33, 4, 1, 1.5, None),      # dg = 4, dilation 1
    (1, 32, 32, 16, 16, 1, 2, 30.0, None),     # one channel block; offsets mostly outside the image
    (1, 32, 32, 24, 52, 2, 2, 1.0, None),      # 16 channels per deformable group (the 1/6 scale): two samples per K block
    (2, 32, 32, 19, 37, 2, 2, 6.0, 2),         # same, ragged tiles, batch 2, many samples outside the staged patch
    (1, 64, 32, 12, 20, 4, 1, 1.5, None),      # 16-channel groups over two channel blocks
])
@pytest.mark.parametrize("rows", [4, 8])
def test_deform_halo_matches_oracle(deform_halo, cfg, rows):
    import aanet_b200.ops as ops
    B, Ci, Co, H, W, dg, dil, sigma, margin = cfg
    os.environ["AANET_DEFORM_ROWS"] = str(rows)          # rows per producer thread: 768- / 512-thread variant
    if margin is None:
        os.environ.pop("AANET_DEFORM_MARGIN", None)
    else:
        os.environ["AANET_DEFORM_MARGIN"] = str(margin)
    rng = np.random.default_rng(23)
    x = rng.standard_normal((B, Ci, H, W)).astype(np.float32)
    off = (sigma * rng.standard_normal((B, dg * 18, H, W))).astype(np.float32)
    msk = (2 / (1 + np.exp(-rng.standard_normal((B, dg * 9, H, W))))).astype(np.float32)
    w = (rng.standard_normal((Co, Ci, 3, 3)) / np.sqrt(Ci * 9)).astype(np.float32)
    sc = (rng.random(Co) + 0.5).astype(np.float32)
    sh = rng.standard_normal(Co).astype(np.float32)
    ref = orc.mdcn_fwd(x, off, msk, w, None, 1, dil, dil, 1, dg)
    ref = np.maximum(ref * sc[None, :, None, None] + sh[None, :, None, None], 0)
    t = [torch.from_numpy(a).cuda() for a in (x, off, msk, w, sc, sh)]
    got = _run(ops, *t, dg, dil)
    assert rel_err(got.cpu().numpy(), ref) < 1e-4
    got_cl = _run(ops, *t, dg, dil, om_nchw=False)               # channels-last offsets/mask
    assert rel_err(got_cl.cpu().numpy(), ref) < 1e-4
    _select(None)
    old = _run(ops, *t, dg, dil)
    _select(deform_halo)
    assert rel_err(got.cpu().numpy(), old.cpu().numpy()) < 2e-5


def test_deform_halo_full_size_vs_gather_engine(deform_halo):
    """The bench's 1/3-scale ISA layer at full size: 2 * randn px offsets (fractional, some outside the image)."""
    import aanet_b200.ops as ops
    torch.manual_seed(326)
    B, C, H, W = 1, 64, 128, 416
    x = torch.randn(B, C, H, W, device="cuda")
    off = 2 * torch.randn(B, 36, H, W, device="cuda")
    msk = 2 * torch.sigmoid(torch.randn(B, 18, H, W, device="cuda"))
    w = torch.randn(C, C, 3, 3, device="cuda") / (C * 9) ** 0.5
    sc, sh = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda")
    new = _run(ops, x, off, msk, w, sc, sh, 2, 2)
    _select(None)
    old = _run(ops, x, off, msk, w, sc, sh, 2, 2)
    _select(deform_halo)
    assert rel_err(new.cpu().numpy(), old.cpu().numpy()) < 2e-5
