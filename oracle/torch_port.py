"""TEST INFRASTRUCTURE ONLY -- CPU port of the module-level hot path.

A functional (state_dict-driven) restatement of the reference's PyTorch CPU path:
cost-volume pyramid -> AdaptiveAggregation (ISA + CSA, 6 modules) -> soft-argmin.
The four custom ops come from the C oracle (oracle/aanet_oracle.c); the dense convolutions
and batch-norms are torch CPU functional calls, which is the same third-party arithmetic the
reference itself runs (SURVEY.md 8c "Third-party arithmetic").

Used by tests/ as the checker for the drop-in modules, and by bench.py for the
`cpu_baseline` / `--impl reference` legs (kind "port").  Never imported by aanet_b200/.

Citations are relative to /root/reference/.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import oracle as orc


def _np(t):
    return t.detach().cpu().contiguous().numpy()


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


# ----------------------------------------------------------------------------- cost volume
def cost_volume_loop(left, right, max_disp):
    """nets/cost.py:40-48 restated with the reference's own per-disparity torch loop (this is
    what the reference's CPU path executes; used as the timed CPU baseline)."""
    b, c, h, w = left.shape
    vol = left.new_zeros(b, max_disp, h, w)
    for d in range(max_disp):
        if d == 0:
            vol[:, 0] = (left * right).mean(1)
        elif d < w:
            vol[:, d, :, d:] = (left[..., d:] * right[..., :-d]).mean(1)
    return vol.contiguous()


def cost_volume_pyramid(lefts, rights, max_disp, use_c=False):
    """nets/cost.py:64-76."""
    out = []
    for s, (l, r) in enumerate(zip(lefts, rights)):
        d = max_disp // (2 ** s)
        out.append(_t(orc.corr_fwd(_np(l), _np(r), d)) if use_c else cost_volume_loop(l, r, d))
    return out


# ----------------------------------------------------------------------------- soft-argmin
def disparity_estimation(cost, match_similarity=True, use_c=True):
    """nets/estimation.py:13-30."""
    if use_c:
        return _t(orc.softargmin_fwd(_np(cost), match_similarity))
    c = cost if match_similarity else -cost
    p = F.softmax(c, dim=1)
    d = torch.arange(c.shape[1], dtype=p.dtype).view(1, -1, 1, 1)
    return (p * d).sum(1)


# ----------------------------------------------------------------------------- ISA pieces
def _bn(x, sd, pre, eps=1e-5):
    """eval-mode nn.BatchNorm2d."""
    return F.batch_norm(x, sd[pre + ".running_mean"], sd[pre + ".running_var"],
                        sd[pre + ".weight"], sd[pre + ".bias"], False, 0.0, eps)


def mdconv(x, offset, mask, weight, bias, stride, pad, dil, groups, dg, impl="c"):
    """nets/deform_conv/deform_conv.py:113-148 -> C oracle, or torchvision (SURVEY 8c secondary
    oracle) when impl == 'tv'."""
    if impl == "tv":
        import torchvision
        return torchvision.ops.deform_conv2d(x, offset, weight, bias, stride=stride, padding=pad,
                                             dilation=dil, mask=mask)
    return _t(orc.mdcn_fwd(_np(x), _np(offset), None if mask is None else _np(mask), _np(weight),
                           None if bias is None else _np(bias), stride, pad, dil, groups, dg))


def deform_conv2d_layer(x, sd, pre, stride=1, dil=2, groups=1, dg=2, k=3, double_mask=True,
                        modulation=True, impl="c"):
    """nets/deform.py:78-97 (DeformConv2d.forward)."""
    om = F.conv2d(x, sd[pre + ".offset_conv.weight"], sd[pre + ".offset_conv.bias"], stride=stride,
                  padding=dil, dilation=dil, groups=dg)
    bias = sd.get(pre + ".deform_conv.bias")
    if not modulation:
        return mdconv(x, om, None, sd[pre + ".deform_conv.weight"], bias, stride, dil, dil, groups,
                      dg, impl)
    n_off = dg * 2 * k * k
    offset = om[:, :n_off].contiguous()
    mask = om[:, n_off:].sigmoid()
    if double_mask:
        mask = mask * 2
    return mdconv(x, offset, mask.contiguous(), sd[pre + ".deform_conv.weight"], bias, stride, dil,
                  dil, groups, dg, impl)


def bottleneck(x, sd, pre, deform, dil=2, dg=2, impl="c"):
    """nets/deform.py:164-184 (SimpleBottleneck) / :216-236 (DeformSimpleBottleneck)."""
    out = F.relu(_bn(F.conv2d(x, sd[pre + ".conv1.weight"]), sd, pre + ".bn1"))
    if deform:
        out = deform_conv2d_layer(out, sd, pre + ".conv2", dil=dil, dg=dg, impl=impl)
    else:
        out = F.conv2d(out, sd[pre + ".conv2.weight"], padding=1)
    out = F.relu(_bn(out, sd, pre + ".bn2"))
    out = _bn(F.conv2d(out, sd[pre + ".conv3.weight"]), sd, pre + ".bn3")
    return F.relu(out + x)


# ----------------------------------------------------------------------------- CSA + module
def _exchange(x, sd, pre, i, j):
    """fuse_layers[i][j] of nets/aggregation.py:346-371 (conv(+BN)(+LeakyReLU) chains)."""
    if i == j:
        return x
    if i < j:
        return _bn(F.conv2d(x, sd[pre + ".0.weight"]), sd, pre + ".1")
    y = x
    for k in range(i - j):
        y = _bn(F.conv2d(y, sd["%s.%d.0.weight" % (pre, k)], stride=2, padding=1), sd,
                "%s.%d.1" % (pre, k))
        if k < i - j - 1:
            y = F.leaky_relu(y, 0.2)
    return y


def aggregation_module(xs, sd, pre, deform, n_out, dil=2, dg=2, impl="c", fuse_c=True):
    """nets/aggregation.py:375-402 (AdaptiveAggregationModule.forward)."""
    xs = [bottleneck(x, sd, "%s.branches.%d.0" % (pre, s), deform, dil, dg, impl)
          for s, x in enumerate(xs)]
    if len(xs) == 1:
        return xs
    outs = []
    for i in range(n_out):
        terms = [_exchange(xs[j], sd, "%s.fuse_layers.%d.%d" % (pre, i, j), i, j)
                 for j in range(len(xs))]
        hw = tuple(terms[0].shape[2:])
        if fuse_c:
            outs.append(_t(orc.csa_fuse_fwd([_np(t) for t in terms], hw, 0.2)))
        else:
            acc = terms[0]
            for t in terms[1:]:
                if tuple(t.shape[2:]) != hw:
                    t = F.interpolate(t, size=hw, mode="bilinear", align_corners=False)
                acc = acc + t
            outs.append(F.leaky_relu(acc, 0.2))
    return outs


def adaptive_aggregation(costs, sd, pre="", num_fusions=6, num_deform_blocks=3,
                         intermediate_supervision=False, dil=2, dg=2, impl="c", fuse_c=True):
    """nets/aggregation.py:452-464 (AdaptiveAggregation.forward); module plan :416-441."""
    xs = list(costs)
    n = len(xs)
    for i in range(num_fusions):
        n_out = n if (intermediate_supervision or i != num_fusions - 1) else 1
        deform = i >= num_fusions - num_deform_blocks
        xs = aggregation_module(xs, sd, "%sfusions.%d" % (pre, i), deform, n_out, dil, dg, impl,
                                fuse_c)
    n_final = n if intermediate_supervision else 1
    return [F.conv2d(xs[s], sd["%sfinal_conv.%d.weight" % (pre, s)],
                     sd["%sfinal_conv.%d.bias" % (pre, s)]) for s in range(n_final)]


def hot_path(lefts, rights, sd, max_disp, agg_prefix="", corr_c=False, impl="c", **kw):
    """nets/aanet.py:216-219: cost pyramid -> aggregation -> soft-argmin (coarse-to-fine list,
    aanet.py:156-167)."""
    costs = cost_volume_pyramid(lefts, rights, max_disp, use_c=corr_c)
    agg = adaptive_aggregation(costs, sd, agg_prefix, impl=impl, **kw)
    return [disparity_estimation(a, True) for a in reversed(agg)]
