"""TMA-staged correlation kernel (aanet_b200/csrc/correlation_tma.cu) against the C oracle (nets/cost.py:40-48
restated) and the golden fixtures made from the reference's CostVolume: MN-major operands straight from NCHW rows,
ragged widths, C not a multiple of 32, every supported window, both output layouts, exact zeros for w < d."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture
def corr_tma():
    old = os.environ.get("AANET_CORR_TMA")
    os.environ["AANET_CORR_TMA"] = "1"
    yield
    if old is None:
        os.environ.pop("AANET_CORR_TMA", None)
    else:
        os.environ["AANET_CORR_TMA"] = old


@pytest.mark.parametrize("shape", [(1, 128, 3, 416, 64), (1, 128, 2, 208, 32), (1, 128, 2, 104, 16),
                                   (2, 32, 4, 200, 64), (1, 5, 3, 36, 7), (1, 33, 2, 140, 20), (3, 8, 2, 260, 48),
                                   (1, 64, 1, 4, 4), (1, 16, 2, 132, 64)])
def test_corr_tma_oracle(corr_tma, shape):
    import aanet_b200.ops as ops
    B, C, H, W, D = shape
    rng = np.random.default_rng(326)
    L = np.maximum(rng.standard_normal((B, C, H, W)), 0).astype(np.float32)
    R = np.maximum(rng.standard_normal((B, C, H, W)), 0).astype(np.float32)
    ref = orc.corr_fwd(L, R, D)
    Lc, Rc = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
    out = ops.correlation(Lc, Rc, D).cpu().numpy()
    assert rel_err(out, ref) < 1e-4
    for d in range(1, D):
        assert np.all(out[:, d, :, :min(d, W)] == 0.0)
    if D % 4 == 0:
        cl = ops.correlation_nhwc(Lc, Rc, D).permute(0, 3, 1, 2).cpu().numpy()
        assert np.array_equal(cl, out)
    os.environ["AANET_CORR_TMA"] = "0"
    old = ops.correlation(Lc, Rc, D).cpu().numpy()
    os.environ["AANET_CORR_TMA"] = "1"
    assert rel_err(out, old) < 1e-5


@pytest.mark.parametrize("tag", ["a", "narrow", "c1", "c128"])
def test_corr_tma_golden(corr_tma, golden, tag):
    import aanet_b200.ops as ops
    z = golden("corr")
    L, R = torch.from_numpy(z[tag + "_L"]).cuda(), torch.from_numpy(z[tag + "_R"]).cuda()
    out = ops.correlation(L, R, int(z[tag + "_D"]))
    assert rel_err(out.cpu().numpy(), z[tag + "_out"]) < 1e-4
