"""Row g1: the drop-in EXECUTED inside the unmodified reference `nets.AANet.forward` (nets/aanet.py:212-229) on
the GPU, against the stock reference (its own modules + its own CUDA op compiled for sm_100a), same weights, same
images, KITTI 384x1248.  Two processes (profiles/full_model.py) because the reference is the top-level package
`nets` in both.  Skipped when the reference was never staged (oracle/stage_ref.py, build container only)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(variant, model, weights, out, extra=()):
    cmd = [sys.executable, os.path.join(ROOT, "profiles", "full_model.py"), "--variant", variant, "--model", model,
           "--weights", weights, "--out", out, "--iters", "3", "--no-graph"] + list(extra)
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    return json.loads(p.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("model,H,W", [("aanet", 384, 1248), ("aanet+", 576, 960)])
def test_dropin_inside_reference_aanet(tmp_path, model, H, W):
    from oracle import build_ref, stage_ref
    if stage_ref.staged_path() is None or build_ref.built_path() is None:
        pytest.skip("reference not staged under baseline/_ref (python oracle/stage_ref.py)")
    w = str(tmp_path / "weights.pt")
    size = ["--height", str(H), "--width", str(W)]
    a = _run("stock", model, w, str(tmp_path / "stock.npz"), size)        # creates the weights
    b = _run("dropin", model, w, str(tmp_path / "dropin.npz"), size)      # loads them strict=True
    assert a["params"] == b["params"] == (3931676 if model == "aanet" else 8442850)
    assert b["aanet_b200_launch_calls"] > 100
    za, zb = np.load(tmp_path / "stock.npz"), np.load(tmp_path / "dropin.npz")
    assert sorted(za.files) == sorted(zb.files) and len(za.files) == 3       # [H/3, H/2, H]
    for k in sorted(za.files):
        assert za[k].shape == zb[k].shape
        err = np.abs(za[k] - zb[k]).max()
        print("%s %s: max |stock - dropin| = %.2e px (range %.1f..%.1f)" % (model, k, err, za[k].min(), za[k].max()))
        assert err < 1e-3, "%s: %g px" % (k, err)
