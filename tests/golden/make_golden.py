"""Generate the golden fixtures in tests/golden/ from the REFERENCE's own Python code.

Run in the build container only (needs /root/reference; it does not exist on the GPU box):

    python tests/golden/make_golden.py

What it imports from the reference (read-only, by path -- nothing is copied):
  * nets/cost.py, nets/estimation.py                     (pure torch)
  * the `nets` package with `nets.deform_conv.deform_conv_cuda` stubbed and
    ModulatedDeformConv/DeformConv.forward routed to torchvision.ops.deform_conv2d
    (the reference op is CUDA-only; torchvision matches its .cu
    semantics bit-for-bit in float64)
and what it writes:  small .npz files with seeded inputs, reference outputs and, where autograd
gives them, reference gradients.  All arrays float32 unless the name ends with `_f64`.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F
import torchvision

REF = os.environ.get("AANET_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))
torch.manual_seed(326)            # the reference's default seed, train.py:44
torch.set_num_threads(4)


def load_by_path(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_nets():
    sys.modules["nets.deform_conv.deform_conv_cuda"] = types.ModuleType("deform_conv_cuda")
    sys.path.insert(0, REF)
    import nets  # noqa
    dc = sys.modules["nets.deform_conv.deform_conv"]
    dc.ModulatedDeformConv.forward = lambda self, x, off, m: torchvision.ops.deform_conv2d(
        x, off, self.weight, self.bias, stride=self.stride, padding=self.padding,
        dilation=self.dilation, mask=m)
    dc.DeformConv.forward = lambda self, x, off: torchvision.ops.deform_conv2d(
        x, off, self.weight, None, stride=self.stride, padding=self.padding,
        dilation=self.dilation)
    return nets


def npf(t):
    return t.detach().cpu().numpy()


def save(name, **arrs):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrs)
    print("%-28s %7.1f KB" % (name + ".npz", os.path.getsize(path) / 1024))


def gen_corr(cost_mod):
    cases = {}
    for tag, (B, C, H, W, D) in {"a": (2, 8, 5, 23, 12), "narrow": (1, 4, 3, 9, 12),
                                 "c1": (1, 1, 2, 16, 4), "c128": (1, 128, 2, 40, 16)}.items():
        L = torch.relu(torch.randn(B, C, H, W)).requires_grad_()
        R = torch.relu(torch.randn(B, C, H, W)).requires_grad_()
        out = cost_mod.CostVolume(D, "correlation")(L, R)
        g = torch.randn_like(out)
        gL, gR = torch.autograd.grad(out, (L, R), g)
        cases.update({tag + "_L": npf(L), tag + "_R": npf(R), tag + "_D": np.int32(D),
                      tag + "_out": npf(out), tag + "_g": npf(g), tag + "_gL": npf(gL),
                      tag + "_gR": npf(gR)})
    # pyramid (cost.py:64-76)
    Ls = [torch.relu(torch.randn(1, 6, 12 >> s, 28 >> s)) for s in range(3)]
    Rs = [torch.relu(torch.randn(1, 6, 12 >> s, 28 >> s)) for s in range(3)]
    pyr = cost_mod.CostVolumePyramid(10, "correlation")(Ls, Rs)
    for s in range(3):
        cases.update({"pyr_L%d" % s: npf(Ls[s]), "pyr_R%d" % s: npf(Rs[s]),
                      "pyr_out%d" % s: npf(pyr[s])})
    cases["pyr_D"] = np.int32(10)
    save("corr", **cases)


def gen_softargmin(est_mod):
    cases = {}
    for tag, (B, D, H, W, sim, scale) in {"sim": (2, 12, 5, 7, True, 1.0),
                                          "cost": (1, 16, 3, 10, False, 3.0),
                                          "d1": (1, 1, 2, 3, True, 1.0),
                                          "peaky": (1, 64, 4, 9, True, 30.0)}.items():
        c = (torch.randn(B, D, H, W) * scale).requires_grad_()
        # max_disp deliberately != D for one case: estimation.py:21-25 uses the tensor's D
        disp = est_mod.DisparityEstimation(D if tag != "cost" else 99, sim)(c)
        g = torch.randn_like(disp)
        gc, = torch.autograd.grad(disp, c, g)
        cases.update({tag + "_cost": npf(c), tag + "_sim": np.int32(sim), tag + "_disp": npf(disp),
                      tag + "_g": npf(g), tag + "_gcost": npf(gc)})
    save("softargmin", **cases)


def gen_mdcn(nets):
    dc = sys.modules["nets.deform_conv.deform_conv"]
    cases = {}
    cfgs = {  # tag: B, Cin, Cout, H, W, k, stride, pad, dil, groups, dg, bias, offset_sigma
        "isa": (2, 8, 8, 9, 14, 3, 1, 2, 2, 1, 2, False, 2.0),
        "s2": (1, 6, 10, 11, 13, 3, 2, 1, 1, 1, 1, True, 2.0),
        "grp": (1, 8, 4, 7, 9, 3, 1, 1, 1, 2, 4, True, 1.0),
        "far": (1, 4, 4, 6, 8, 3, 1, 2, 2, 1, 2, False, 6.0),
        "k1": (1, 4, 6, 5, 7, 1, 1, 0, 1, 1, 1, False, 1.5),
    }
    for tag, (B, Ci, Co, H, W, k, st, pad, dil, grp, dg, bias, sig) in cfgs.items():
        m = dc.ModulatedDeformConv(Ci, Co, k, stride=st, padding=pad, dilation=dil, groups=grp,
                                   deformable_groups=dg, bias=bias).double()
        if bias:
            torch.nn.init.normal_(m.bias, std=0.5)
        if grp != 1:
            # torchvision infers groups from the weight shape; same semantics as cpp:544-555
            pass
        Ho = (H + 2 * pad - (dil * (k - 1) + 1)) // st + 1
        Wo = (W + 2 * pad - (dil * (k - 1) + 1)) // st + 1
        x = torch.randn(B, Ci, H, W, dtype=torch.float64).requires_grad_()
        off = (sig * torch.randn(B, dg * 2 * k * k, Ho, Wo, dtype=torch.float64)).requires_grad_()
        msk = (2 * torch.sigmoid(torch.randn(B, dg * k * k, Ho, Wo, dtype=torch.float64))
               ).requires_grad_()
        out = m(x, off, msk)
        g = torch.randn_like(out)
        params = (x, off, msk, m.weight) + ((m.bias,) if bias else ())
        grads = torch.autograd.grad(out, params, g)
        cases.update({tag + "_cfg": np.array([st, pad, dil, grp, dg, int(bias)], np.int32),
                      tag + "_x_f64": npf(x), tag + "_off_f64": npf(off), tag + "_mask_f64": npf(msk),
                      tag + "_w_f64": npf(m.weight), tag + "_out_f64": npf(out), tag + "_g_f64": npf(g),
                      tag + "_gx_f64": npf(grads[0]), tag + "_goff_f64": npf(grads[1]),
                      tag + "_gmask_f64": npf(grads[2]), tag + "_gw_f64": npf(grads[3])})
        if bias:
            cases.update({tag + "_b_f64": npf(m.bias), tag + "_gb_f64": npf(grads[4])})
    # DCNv1 (deform_conv.py:190-239): no mask
    m = dc.DeformConv(6, 4, 3, stride=1, padding=1, dilation=1, groups=1, deformable_groups=2).double()
    x = torch.randn(2, 6, 7, 8, dtype=torch.float64).requires_grad_()
    off = (1.5 * torch.randn(2, 2 * 18, 7, 8, dtype=torch.float64)).requires_grad_()
    out = m(x, off)
    g = torch.randn_like(out)
    gx, goff, gw = torch.autograd.grad(out, (x, off, m.weight), g)
    cases.update({"v1_cfg": np.array([1, 1, 1, 1, 2, 0], np.int32), "v1_x_f64": npf(x),
                  "v1_off_f64": npf(off), "v1_w_f64": npf(m.weight), "v1_out_f64": npf(out),
                  "v1_g_f64": npf(g), "v1_gx_f64": npf(gx), "v1_goff_f64": npf(goff),
                  "v1_gw_f64": npf(gw)})
    save("mdcn", **cases)


def gen_deformconv2d(nets):
    """DeformConv2d (nets/deform.py:17-97) incl. the positional offset/mask slice quirk."""
    from nets.deform import DeformConv2d, DeformSimpleBottleneck
    cases = {}
    layer = DeformConv2d(8, 8, dilation=2, deformable_groups=2)
    torch.nn.init.normal_(layer.offset_conv.weight, std=0.3)
    torch.nn.init.normal_(layer.offset_conv.bias, std=0.5)
    x = torch.randn(2, 8, 10, 12)
    cases["dc2d_x"] = npf(x)
    cases["dc2d_out"] = npf(layer(x))
    for k, v in layer.state_dict().items():
        cases["dc2d_sd/" + k] = npf(v)
    blk = DeformSimpleBottleneck(8, 8, mdconv_dilation=2, deformable_groups=2).eval()
    randomize(blk)
    cases["blk_x"] = npf(x)
    cases["blk_out"] = npf(blk(x))
    for k, v in blk.state_dict().items():
        cases["blk_sd/" + k] = npf(v)
    save("deform_layer", **cases)


def randomize(mod):
    """Non-trivial BN statistics and offset_conv weights (both are identity-like at init:
    deform.py:75-76, default BN), otherwise the bilinear path is never exercised."""
    for n, m in mod.named_modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.2)
            m.running_var.uniform_(0.5, 1.5)
            torch.nn.init.uniform_(m.weight, 0.5, 1.5)
            torch.nn.init.normal_(m.bias, std=0.2)
        if n.endswith("offset_conv"):
            torch.nn.init.normal_(m.weight, std=0.15)
            torch.nn.init.normal_(m.bias, std=0.5)


def gen_csa():
    """aggregation.py:387-400 restated with F.interpolate directly: sizes incl. odd ones."""
    cases = {}
    shapes = {"x2x4": ((2, 3, 12, 20), [(12, 20), (6, 10), (3, 5)]),
              "odd": ((1, 2, 9, 13), [(9, 13), (5, 7), (3, 4)]),
              "same": ((1, 4, 5, 6), [(5, 6), (5, 6), (5, 6)]),
              "two": ((1, 2, 8, 8), [(8, 8), (4, 4)])}
    for tag, ((B, C, H, W), ths) in shapes.items():
        terms = [torch.randn(B, C, h, w).requires_grad_() for (h, w) in ths]
        acc = terms[0]
        for t in terms[1:]:
            if t.shape[2:] != acc.shape[2:]:
                t = F.interpolate(t, size=acc.shape[2:], mode="bilinear", align_corners=False)
            acc = acc + t
        out = F.leaky_relu(acc, 0.2)
        g = torch.randn_like(out)
        grads = torch.autograd.grad(out, terms, g)
        cases[tag + "_n"] = np.int32(len(terms))
        cases[tag + "_out"] = npf(out)
        cases[tag + "_g"] = npf(g)
        for i, (t, gt) in enumerate(zip(terms, grads)):
            cases["%s_t%d" % (tag, i)] = npf(t)
            cases["%s_gt%d" % (tag, i)] = npf(gt)
    save("csa", **cases)


def gen_aggregation(nets, cost_mod, est_mod, cases_=(("agg", False, 16, 24, 36), ("agg_inter", True, 16, 24, 36))):
    """End-to-end hot path on a miniature pyramid: CostVolumePyramid -> AdaptiveAggregation
    (6 modules, last 3 deformable, as nets/aanet.py:31,92-99 builds it) -> DisparityEstimation.
    `agg32` (D0 = 32: 32/16/8 channels, every count a multiple of 4 * deformable_groups) is the fixture that
    reaches the channels-last tcgen05 executor (aanet_b200/fused.py: supported())."""
    from nets.aggregation import AdaptiveAggregation
    for tag, inter, D0, H, W in cases_:
        C = 8
        agg = AdaptiveAggregation(max_disp=D0, num_scales=3, num_fusions=6, num_stage_blocks=1,
                                  num_deform_blocks=3, intermediate_supervision=inter,
                                  deformable_groups=2, mdconv_dilation=2).eval()
        randomize(agg)
        Ls = [torch.relu(torch.randn(1, C, H >> s, W >> s)) for s in range(3)]
        Rs = [torch.relu(torch.randn(1, C, H >> s, W >> s)) for s in range(3)]
        with torch.no_grad():
            costs = cost_mod.CostVolumePyramid(D0, "correlation")(Ls, Rs)
            outs = agg([c.clone() for c in costs])
            disps = [est_mod.DisparityEstimation(D0, True)(o) for o in reversed(outs)]
        cases = {"D0": np.int32(D0), "inter": np.int32(inter)}
        for s in range(3):
            cases["L%d" % s] = npf(Ls[s]); cases["R%d" % s] = npf(Rs[s])
            cases["cost%d" % s] = npf(costs[s])
        for i, o in enumerate(outs):
            cases["agg%d" % i] = npf(o)
        for i, d in enumerate(disps):
            cases["disp%d" % i] = npf(d)
        for k, v in agg.state_dict().items():
            if not k.endswith("num_batches_tracked"):
                cases["sd/" + k] = npf(v)
        save(tag, **cases)


def gen_refinement(nets):
    """Refinement: the front end of StereoDRNetRefinement / HourglassRefinement.forward
    (refinement.py:80-95, :144-160 -> warp.py:41-64) captured with forward hooks on the reference modules
    themselves -- conv1 sees cat(error, left), conv2 sees the upsampled, rescaled disparity -- plus the whole
    StereoDRNetRefinement output and state_dict (the hourglass weights would be 9 MB: its key/shape parity is
    checked against the live reference in tests/test_host.py instead)."""
    torch.manual_seed(327)
    from nets.refinement import StereoDRNetRefinement
    cases = {}
    net = StereoDRNetRefinement().eval()
    randomize(net)
    seen = {}
    net.conv1.register_forward_hook(lambda m, i, o: seen.__setitem__("concat", i[0].clone()))
    net.conv2.register_forward_hook(lambda m, i, o: seen.__setitem__("disp", i[0].clone()))
    shapes = {"x3": (2, 8, 12, 24, 36), "same": (1, 10, 14, 10, 14), "odd": (1, 5, 7, 13, 20),
              "x2": (1, 16, 24, 32, 48)}
    for tag, (B, h, w, H, W) in shapes.items():
        low = torch.rand(B, h, w) * (w / 2.5)            # positive; x - disp leaves the image on the left
        left, right = torch.rand(B, 3, H, W), torch.rand(B, 3, H, W)
        with torch.no_grad():
            out = net(low, left, right)
        cases.update({tag + "_low": npf(low), tag + "_left": npf(left), tag + "_right": npf(right),
                      tag + "_concat": npf(seen["concat"]), tag + "_disp": npf(seen["disp"]),
                      tag + "_drnet_out": npf(out)})
    for k, v in net.state_dict().items():
        if not k.endswith("num_batches_tracked"):
            cases["drnet_sd/" + k] = npf(v)
    save("refinement", **cases)


if __name__ == "__main__":
    cost_mod = load_by_path("ref_cost", "nets/cost.py")
    est_mod = load_by_path("ref_estimation", "nets/estimation.py")
    nets = load_nets()
    if "--only-refinement" in sys.argv:     # added after the other fixtures were committed; own seed
        gen_refinement(nets)
        sys.exit(0)
    if "--only-cost5d" in sys.argv:         # round 2: the 5-D 'difference' / 'concat' volumes (cost.py:22-38)
        torch.manual_seed(329)
        cases = {}
        for kind in ("difference", "concat"):
            for tag, (B, C, H, W, D) in {"a": (2, 3, 4, 13, 6), "narrow": (1, 2, 2, 5, 8)}.items():
                L = torch.randn(B, C, H, W).requires_grad_(); R = torch.randn(B, C, H, W).requires_grad_()
                out = cost_mod.CostVolume(D, kind)(L, R)
                g = torch.randn_like(out)
                gL, gR = torch.autograd.grad(out, (L, R), g)
                k = "%s_%s_" % (kind, tag)
                cases.update({k + "L": npf(L), k + "R": npf(R), k + "D": np.int32(D), k + "out": npf(out),
                              k + "g": npf(g), k + "gL": npf(gL), k + "gR": npf(gR)})
        save("cost5d", **cases)
        sys.exit(0)
    if "--only-agg32" in sys.argv:          # round 2: a reference-made fixture that reaches the fused executor
        torch.manual_seed(328)
        gen_aggregation(nets, cost_mod, est_mod, (("agg32", False, 32, 20, 28),))
        sys.exit(0)
    gen_corr(cost_mod)
    gen_softargmin(est_mod)
    gen_mdcn(nets)
    gen_deformconv2d(nets)
    gen_csa()
    gen_aggregation(nets, cost_mod, est_mod)
    gen_refinement(nets)
