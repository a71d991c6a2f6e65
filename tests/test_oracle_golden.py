"""Pin the CPU oracle (oracle/) against fixtures produced by the reference's own Python code
(tests/golden/make_golden.py).  Runs without a GPU."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import oracle as orc
from oracle import torch_port as port


@pytest.mark.parametrize("tag", ["a", "narrow", "c1", "c128"])
def test_corr_fwd_bwd(golden, tag):
    z = golden("corr")
    out = orc.corr_fwd(z[tag + "_L"], z[tag + "_R"], int(z[tag + "_D"]))
    assert out.shape == z[tag + "_out"].shape
    assert rel_err(out, z[tag + "_out"]) < 2e-6
    gL, gR = orc.corr_bwd(z[tag + "_L"], z[tag + "_R"], z[tag + "_g"])
    assert rel_err(gL, z[tag + "_gL"]) < 2e-6
    assert rel_err(gR, z[tag + "_gR"]) < 2e-6


def test_corr_zero_band_and_pyramid(golden):
    z = golden("corr")
    out = orc.corr_fwd(z["narrow_L"], z["narrow_R"], int(z["narrow_D"]))   # W=9 < D=12
    for d in range(out.shape[1]):
        assert np.all(out[:, d, :, :d] == 0.0)
    pyr = orc.corr_pyramid([z["pyr_L%d" % s] for s in range(3)],
                           [z["pyr_R%d" % s] for s in range(3)], int(z["pyr_D"]))
    for s in range(3):
        assert pyr[s].shape == z["pyr_out%d" % s].shape
        assert rel_err(pyr[s], z["pyr_out%d" % s]) < 2e-6
    # the timed CPU-baseline restatement (torch loop) agrees with the C one
    t = port.cost_volume_loop(torch.from_numpy(z["a_L"]), torch.from_numpy(z["a_R"]), int(z["a_D"]))
    assert rel_err(t.numpy(), z["a_out"]) < 1e-6


@pytest.mark.parametrize("tag", ["sim", "cost", "d1", "peaky"])
def test_softargmin(golden, tag):
    z = golden("softargmin")
    sim = bool(z[tag + "_sim"])
    disp = orc.softargmin_fwd(z[tag + "_cost"], sim)
    assert np.abs(disp - z[tag + "_disp"]).max() < 2e-5          # px
    gc = orc.softargmin_bwd(z[tag + "_cost"], z[tag + "_g"], sim)
    assert rel_err(gc, z[tag + "_gcost"]) < 1e-5


@pytest.mark.parametrize("tag", ["isa", "s2", "grp", "far", "k1", "v1"])
def test_mdcn_f64(golden, tag):
    z = golden("mdcn")
    st, pad, dil, grp, dg, has_b = [int(v) for v in z[tag + "_cfg"]]
    mask = z.get(tag + "_mask_f64")
    bias = z.get(tag + "_b_f64")
    out = orc.mdcn_fwd(z[tag + "_x_f64"], z[tag + "_off_f64"], mask, z[tag + "_w_f64"], bias,
                       st, pad, dil, grp, dg)
    assert rel_err(out, z[tag + "_out_f64"]) < 1e-12
    gx, goff, gmask, gw, gb = orc.mdcn_bwd(z[tag + "_x_f64"], z[tag + "_off_f64"], mask,
                                           z[tag + "_w_f64"], z[tag + "_g_f64"], bool(has_b),
                                           st, pad, dil, grp, dg)
    assert rel_err(gx, z[tag + "_gx_f64"]) < 1e-12
    assert rel_err(goff, z[tag + "_goff_f64"]) < 1e-12
    assert rel_err(gw, z[tag + "_gw_f64"]) < 1e-12
    if mask is not None:
        assert rel_err(gmask, z[tag + "_gmask_f64"]) < 1e-12
    if has_b:
        assert rel_err(gb, z[tag + "_gb_f64"]) < 1e-12


def test_mdcn_f32_close_to_f64(golden):
    z = golden("mdcn")
    f = lambda k: z[k].astype(np.float32)
    out = orc.mdcn_fwd(f("isa_x_f64"), f("isa_off_f64"), f("isa_mask_f64"), f("isa_w_f64"), None,
                       1, 2, 2, 1, 2)
    assert out.dtype == np.float32
    assert rel_err(out, z["isa_out_f64"]) < 5e-6


@pytest.mark.parametrize("tag", ["x2x4", "odd", "same", "two"])
def test_csa_fuse(golden, tag):
    z = golden("csa")
    n = int(z[tag + "_n"])
    terms = [z["%s_t%d" % (tag, i)] for i in range(n)]
    out = orc.csa_fuse_fwd(terms, z[tag + "_out"].shape[2:], 0.2)
    assert rel_err(out, z[tag + "_out"]) < 1e-6
    gts = orc.csa_fuse_bwd(out, z[tag + "_g"], [t.shape[2:] for t in terms], 0.2)
    for i in range(n):
        assert rel_err(gts[i], z["%s_gt%d" % (tag, i)]) < 1e-6


def _sd(z, pre):
    return {k[len(pre):]: torch.from_numpy(v) for k, v in z.items() if k.startswith(pre)}


def test_deform_layer_port(golden):
    z = golden("deform_layer")
    x = torch.from_numpy(z["dc2d_x"])
    sd = {"conv2." + k: v for k, v in _sd(z, "dc2d_sd/").items()}
    out = port.deform_conv2d_layer(x, sd, "conv2", dil=2, dg=2)
    assert rel_err(out.numpy(), z["dc2d_out"]) < 1e-5
    blk = port.bottleneck(torch.from_numpy(z["blk_x"]), {"b." + k: v for k, v in _sd(z, "blk_sd/").items()},
                          "b", deform=True)
    assert rel_err(blk.numpy(), z["blk_out"]) < 1e-5


@pytest.mark.parametrize("name", ["agg", "agg_inter", "agg32"])
def test_hot_path_port(golden, name):
    z = golden(name)
    sd = _sd(z, "sd/")
    inter = bool(z["inter"])
    Ls = [torch.from_numpy(z["L%d" % s]) for s in range(3)]
    Rs = [torch.from_numpy(z["R%d" % s]) for s in range(3)]
    costs = port.cost_volume_pyramid(Ls, Rs, int(z["D0"]), use_c=True)
    for s in range(3):
        assert rel_err(costs[s].numpy(), z["cost%d" % s]) < 2e-6
    outs = port.adaptive_aggregation(costs, sd, intermediate_supervision=inter)
    assert len(outs) == (3 if inter else 1)
    for i, o in enumerate(outs):
        assert rel_err(o.numpy(), z["agg%d" % i]) < 1e-4
    disps = [port.disparity_estimation(o) for o in reversed(outs)]
    for i, d in enumerate(disps):
        assert np.abs(d.numpy() - z["disp%d" % i]).max() < 1e-3


# ------------------------------------------------------------------------------------ refinement front end
@pytest.mark.parametrize("tag", ["x3", "same", "odd", "x2"])
def test_refine_frontend_oracle_vs_reference(golden, tag):
    """Oracle vs tensors captured inside the reference's StereoDRNetRefinement.forward (inputs of conv1 / conv2)."""
    z = golden("refinement")
    concat, disp = orc.refine_frontend_fwd(z[tag + "_low"], z[tag + "_left"], z[tag + "_right"])
    assert np.abs(disp - z[tag + "_disp"]).max() < 1e-5 * max(1.0, np.abs(z[tag + "_disp"]).max())
    assert np.abs(concat - z[tag + "_concat"]).max() < 1e-5
    assert np.array_equal(concat[:, 3:], z[tag + "_left"])


@pytest.mark.parametrize("kind", ["difference", "concat"])
@pytest.mark.parametrize("tag", ["a", "narrow"])
def test_cost5d_oracle_golden(golden, kind, tag):
    """numpy restatement of nets/cost.py:22-38 against the reference's own CostVolume (bit-exact forward)."""
    from oracle import oracle as orc
    z = golden("cost5d")
    k = "%s_%s_" % (kind, tag)
    out = orc.cost5d_fwd(z[k + "L"], z[k + "R"], int(z[k + "D"]), kind)
    assert out.shape == z[k + "out"].shape and np.array_equal(out, z[k + "out"])
    gL, gR = orc.cost5d_bwd(z[k + "g"], kind)
    assert rel_err(gL, z[k + "gL"]) < 1e-6 and rel_err(gR, z[k + "gR"]) < 1e-6
