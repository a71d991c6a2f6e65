"""Fused, channels-last inference executor for AdaptiveAggregation.

Same arithmetic as AdaptiveAggregation.forward in eval mode (reference nets/aggregation.py:452-464 and
the modules below it), but every convolution -- the 1x1 / 3x3 convs of the bottlenecks (deform.py:164-184,
:216-236), the offset/mask head (deform.py:80-89), the CSA exchange convs (aggregation.py:346-371) and the
final 1x1 (aggregation.py:443-450) -- runs on the tcgen05 kernels (TMEM-A kernel deform_tmem.cu where the layer has
32-channel blocks, round-1 engine otherwise) with eval-mode BatchNorm folded into a per-channel (scale, shift)
epilogue, ReLU / LeakyReLU / residual add fused, and activations kept channels-last between kernels.  Launches per
stereo pair at config 2: 96 (module-by-module path: ~320).  Fusions beyond single layers: conv3 inside conv2's
launch (_Bottleneck), the coarsest CSA row's sum inside its exchange convolutions (go_folded), the last module's
sum + final 1x1 + soft-argmin as one launch (to_disp).

This is SURVEY.md 8(f) ranks 1-2.  It is used automatically by AdaptiveAggregation.forward when the module
is in eval mode, autograd is off and every channel count is a multiple of 4; otherwise the module-by-module
path (torch convs + the same sm_100a operators) runs.
"""
import os

import torch
import torch.nn as nn

from . import ops
from .streams import fork_join
from .nets.deform import DeformSimpleBottleneck, SimpleBottleneck, bn_affine


def _own(t):
    """Private fp32 copy whose address stays valid for the executor's lifetime (see refresh())."""
    return None if t is None else t.detach().float().clone().contiguous()


def _affine(bn):
    return (None, None) if bn is None else tuple(_own(t) for t in bn_affine(bn))


class _Conv:
    """A convolution prepared for the engine: packed weights + epilogue vectors + geometry.

    The packed weights and the folded-BN vectors are PRIVATE buffers; CUDA graphs captured over the executor
    (HotPath.capture, HostPipeline) bake their addresses.  refresh() therefore rewrites them in place after a
    weight update (load_state_dict, optimizer step) instead of allocating new ones, so a replay of an earlier
    capture sees the new weights -- the same contract as a captured torch module."""

    def __init__(self, conv, bn=None, act=ops.ACT_NONE, slope=0.2):
        w = conv.weight
        self.Cout, _, self.kh, self.kw = w.shape
        self.groups = conv.groups
        self.stride, self.pad, self.dil = conv.stride[0], conv.padding[0], conv.dilation[0]
        self.wpack = ops.pack_conv_weight(w, conv.groups)
        self._conv, self._bn, self._w, self._packs = conv, bn, w, {}
        self.bias = _own(conv.bias)
        self.scale, self.shift = _affine(bn)
        self.act, self.slope = act, slope

    def refresh(self):
        self._w = self._conv.weight
        ops.pack_conv_weight(self._w, self.groups, out=self.wpack)
        for bn, pack in self._packs.items():
            ops.pack_conv_weight(self._w, self.groups, bn, out=pack)
        if self.bias is not None:
            self.bias.copy_(self._conv.bias.detach())
        if self._bn is not None:
            sc, sh = bn_affine(self._bn)
            self.scale.copy_(sc); self.shift.copy_(sh)

    def problem(self, x, bn):
        """Descriptor of this conv for a multi-problem launch whose N tile is `bn` wide."""
        if bn not in self._packs:
            self._packs[bn] = ops.pack_conv_weight(self._w, self.groups, bn)
        return dict(x=x, wpack=self._packs[bn], Cout=self.Cout, kh=self.kh, kw=self.kw, bias=self.bias,
                    scale=self.scale, shift=self.shift, act=self.act, slope=self.slope, stride=self.stride,
                    pad=self.pad, dil=self.dil, groups=self.groups)

    def tmem_eligible(self):
        """Multi-tap convolution over 32-channel blocks (stride 1 or 2): runs as its own launch of the TMEM-A kernel
        (TMA-staged input patch, no per-K-block L2 round trips) instead of inside a multi-problem engine launch."""
        if os.environ.get("AANET_DENSE_TMEM", "1") == "0" or os.environ.get("AANET_EXCHANGE_TMEM", "1") == "0":
            return False
        if self.kh * self.kw == 1:       # 1x1 exchange convs as own TMEM-A launches: measured neutral (1194 vs 1191 pairs/s), opt-in
            return (self.stride == 1 and self.pad == 0 and (self._w.shape[1] % 32) == 0 and self.groups == 1
                    and os.environ.get("AANET_POINTWISE_TMEM", "0") == "1")
        return self.kh * self.kw >= 3 and self.stride in (1, 2) and (self._w.shape[1] % 32) == 0

    def single(self, x, residual=None, act=None, slope=None):
        """Own launch; the N tile is at least 32 wide so that narrow outputs (the 16-channel scale) qualify too."""
        bn = max(32, ops.natural_bn(self.Cout // self.groups))
        q = self.problem(x, bn)
        if residual is not None:
            q["residual"] = residual
        if act is not None:
            q["act"] = act
        if slope is not None:
            q["slope"] = slope
        return ops.conv_batch([q], bn=bn)[0]

    def as_tail(self, residual, act=None):
        """This (1x1) convolution as the fused tail of the preceding one (ops.conv_batch "tail")."""
        return dict(wpack=self.wpack, Cout=self.Cout, scale=self.scale, shift=self.shift, residual=residual,
                    act=self.act if act is None else act)

    def __call__(self, x, residual=None, out_nchw=False, n_offset_ch=0, mask_scale=1.0, act=None, tail=None):
        return ops.conv2d_nhwc(x, self.wpack, self.Cout, self.kh, self.kw, self.bias, self.scale, self.shift,
                               residual, self.act if act is None else act, self.slope, self.stride, self.pad,
                               self.dil, self.groups, out_nchw, n_offset_ch, mask_scale, tail=tail)


class _Deform:
    """DeformConv2d (offset/mask head + DCNv2) with bn2 + ReLU folded into the DCN epilogue."""

    def __init__(self, dc2d, bn):
        dc = dc2d.deform_conv
        self.head = _Conv(dc2d.offset_conv)
        self.n_off = dc2d.deformable_groups * 2 * dc2d.kernel_size * dc2d.kernel_size
        self.mask_scale = 2.0 if dc2d.double_mask else 1.0
        self.Cout, _, self.kh, self.kw = dc.weight.shape
        self.wpack = ops.pack_conv_weight(dc.weight, dc.groups)
        self._dc, self._bn = dc, bn
        self.bias = _own(dc.bias)
        self.scale, self.shift = _affine(bn)
        self.stride, self.pad, self.dil = dc.stride, dc.padding, dc.dilation
        self.groups, self.dg = dc.groups, dc.deformable_groups

    def refresh(self):
        self.head.refresh()
        ops.pack_conv_weight(self._dc.weight, self.groups, out=self.wpack)
        if self.bias is not None:
            self.bias.copy_(self._dc.bias.detach())
        sc, sh = bn_affine(self._bn)
        self.scale.copy_(sc); self.shift.copy_(sh)

    def __call__(self, x, tail=None):
        # offsets pass through, mask channels get mask_scale * sigmoid (deform.py:82-89), one tensor, written
        # as channel planes: the DCN producers read it pixel-contiguously
        om = self.head(x, act=ops.ACT_OFFSET_MASK, n_offset_ch=self.n_off, mask_scale=self.mask_scale,
                       out_nchw=True)
        return ops.mdcn_nhwc(x, om, self.wpack, self.Cout, self.kh, self.kw, self.bias, self.scale, self.shift,
                             True, self.stride, self.pad, self.dil, self.groups, self.dg, om_nchw=True, tail=tail)


class _Bottleneck:
    def __init__(self, blk):
        self.c1 = _Conv(blk.conv1, blk.bn1, ops.ACT_RELU)
        if isinstance(blk, DeformSimpleBottleneck):
            self.c2 = _Deform(blk.conv2, blk.bn2)
        else:
            self.c2 = _Conv(blk.conv2, blk.bn2, ops.ACT_RELU)
        self.c3 = _Conv(blk.conv3, blk.bn3, ops.ACT_RELU)     # relu(bn3(conv3) + identity)
        self._tail = {}

    def refresh(self):
        for c in (self.c1, self.c2, self.c3):
            c.refresh()

    def _tail_ok(self, y1):
        """Is conv3 (1x1 + bn3 + identity + ReLU) fusable into conv2's launch for this input shape?  (Tensor-memory
        kernels: 32-channel blocks, one N tile; asked once per shape.)"""
        # the A/B switches of the kernels are read per launch, so they are part of the key
        key = (tuple(y1.shape), y1.device) + tuple(os.environ.get(k) for k in
                                                   ("AANET_DENSE_TMEM", "AANET_DEFORM_TMEM", "AANET_TAIL_FUSION"))
        if key not in self._tail:
            c2, c3 = self.c2, self.c3
            ok = c3.kh == 1 and c3.kw == 1 and c3.groups == 1 and c3.bias is None and c3.stride == 1 and c3.pad == 0
            if ok:
                deform = isinstance(c2, _Deform)
                q = dict(x=y1, wpack=c2.wpack, Cout=c2.Cout, kh=c2.kh, kw=c2.kw, bias=c2.bias, scale=c2.scale,
                         shift=c2.shift, act=ops.ACT_RELU, stride=c2.stride, pad=c2.pad, dil=c2.dil, groups=c2.groups,
                         tail=c3.as_tail(None))
                if deform:
                    B, H, W, _ = y1.shape
                    q.update(dg=c2.dg, om_nchw=True, offmask=y1.new_empty(B, c2.n_off * 3 // 2, H, W))
                ok = ops.conv_tail_supported(q, deform)
            self._tail[key] = ok
        return self._tail[key]

    def conv1_fusable(self):
        """conv1 as the consumer of the fused CSA sum (ops.csa_conv1_nhwc): plain 1x1 + bn1 + ReLU."""
        c = self.c1
        return (c.kh == 1 and c.kw == 1 and c.stride == 1 and c.pad == 0 and c.groups == 1 and c.bias is None
                and c.act == ops.ACT_RELU and c.Cout in (32, 64))

    def __call__(self, x, y1=None):
        if y1 is None:                      # (else: conv1 was computed by the CSA launch that produced x)
            y1 = self.c1(x)
        if self._tail_ok(y1):
            return self.c2(y1, tail=self.c3.as_tail(x))          # conv2 (+ bn2 + ReLU) and conv3 + bn3 + x + ReLU: one launch
        return self.c3(self.c2(y1), residual=x)


# One multi-problem launch for the last conv of all exchange chains of a CSA row (A/B on one box: 913 vs 910
# pairs/s, 10 launches fewer per pair).  AANET_BATCH_EXCHANGE=0 restores one launch per conv.
BATCH_EXCHANGE = os.environ.get("AANET_BATCH_EXCHANGE", "1") == "1"
# Coarsest CSA row without a csa_fuse launch (sum folded into the exchange convolutions' epilogues); =0: A/B switch
FOLD_LAST_ROW = os.environ.get("AANET_FOLD_LAST_ROW", "1") == "1"
# last module: sum + final 1x1 + soft-argmin as one launch (ops.csa_conv1_nhwc, ACT_SOFTARGMIN); =0: A/B switch
FUSE_FINAL = os.environ.get("AANET_FUSE_FINAL", "1") == "1"
# opt-in: exchange convolutions that only need a coarse scale's output run inside that scale's ISA branch
EARLY_EXCHANGE = os.environ.get("AANET_EARLY_EXCHANGE", "0") == "1"     # measured slower (1188 vs 1231 pairs/s): SM-time is the currency
FOLD_PARALLEL = os.environ.get("AANET_FOLD_PARALLEL", "0") == "1"      # measured slower (1199 vs 1222 pairs/s)


def _exchange(seq):
    """fuse_layers[i][j]: Identity, Sequential(conv, bn) or a chain of Sequential(conv, bn[, LeakyReLU])."""
    if isinstance(seq, nn.Identity):
        return []
    if isinstance(seq[0], nn.Conv2d):
        return [_Conv(seq[0], seq[1])]
    return [_Conv(s[0], s[1], ops.ACT_LEAKY if len(s) > 2 else ops.ACT_NONE,
                  s[2].negative_slope if len(s) > 2 else 0.2) for s in seq]


def supported(agg):
    """All engine constraints: plain BatchNorm2d, every channel count % 4 == 0, modulated deform convs."""
    for m in agg.modules():
        if isinstance(m, nn.Conv2d) and (m.in_channels % 4 or (m.in_channels // m.groups) % 4):
            return False
        if isinstance(m, (SimpleBottleneck, DeformSimpleBottleneck)):
            if m.downsample is not None or not isinstance(m.bn1, nn.BatchNorm2d):
                return False
        if isinstance(m, DeformSimpleBottleneck):
            dc = m.conv2.deform_conv
            if not m.conv2.modulation or dc.in_channels % (4 * dc.deformable_groups) or \
                    (dc.in_channels // dc.groups) % 4:
                return False
    return True


def _state_key(agg):
    return tuple((t.data_ptr(), t._version) for t in list(agg.parameters()) + list(agg.buffers()))


def _shape_key(agg):
    return tuple((tuple(t.shape), t.device) for t in list(agg.parameters()) + list(agg.buffers()))


class FusedAggregation:
    def __init__(self, agg):
        self.key, self.shapes = _state_key(agg), _shape_key(agg)
        self.stages = []
        for mod in agg.fusions:
            branches = [[_Bottleneck(b) for b in br] for br in mod.branches]
            fuse = [[_exchange(mod.fuse_layers[i][j]) for j in range(mod.num_scales)]
                    for i in range(len(mod.fuse_layers))] if mod.num_scales > 1 else None
            self.stages.append((branches, fuse, mod.relu.negative_slope))
        self.final = [_Conv(c) for c in agg.final_conv]

    def refresh(self, agg):
        """Re-pack every layer into the buffers it already owns (same addresses) after a weight update."""
        for branches, fuse, _ in self.stages:
            for br in branches:
                for blk in br:
                    blk.refresh()
            for row in fuse or []:
                for chain in row:
                    for conv in chain:
                        conv.refresh()
        for conv in self.final:
            conv.refresh()
        self.key = _state_key(agg)

    def disparity_fusable(self):
        """Final 1x1 convolution + soft-argmin as one launch (ops.ACT_SOFTARGMIN): plain 1x1 convs whose disparity
        candidates fit one N tile.  AANET_FUSE_SOFTARGMIN=0: A/B switch."""
        return (os.environ.get("AANET_FUSE_SOFTARGMIN", "1") != "0" and
                all(c.kh == 1 and c.kw == 1 and c.groups == 1 and c.Cout <= 64 and c.stride == 1 and c.pad == 0
                    for c in self.final))

    def __call__(self, cost_volume, nhwc=False, disparity=False):
        """cost_volume: the pyramid of volumes, [B,D,H,W] each -- or already channels-last [B,H,W,D] (nhwc=True,
        ops.correlation_nhwc), which saves the three layout kernels.

        Memory plan.  Lifetimes follow the stage structure: inside a branch or a CSA row every intermediate is
        produced and consumed on ONE stream, so it is released as soon as its Python reference dies (the caching
        allocator's stream-ordered reuse is then safe); the only tensors that cross streams are a stage's outputs,
        and those stay referenced (`xs`) until the join that ends the NEXT stage, after which any reuse is ordered
        behind that join by the following fork (streams.py).  Peak: the stage inputs + ~3 volumes per scale, i.e.
        ~6 volumes of the 1/3-scale size per pair instead of the ~70 a keep-everything plan holds."""
        dev = cost_volume[0].device

        def folded(i, row):
            """Row i of a CSA stage whose sum is folded into the exchange convolutions' epilogues (see go_folded)."""
            return (FOLD_LAST_ROW and i == len(row) - 1 and i > 0 and not row[i] and
                    all(c and c[-1].act == ops.ACT_NONE and c[-1].bias is None for j, c in enumerate(row) if j != i))

        def run_chain(chain, x, **last):
            for conv in chain[:-1]:
                x = conv.single(x) if conv.tmem_eligible() else conv(x)
            c = chain[-1]
            return c.single(x, **last) if (last or c.tmem_eligible()) else c(x)

        def branch(s, blocks, x, y1, fuse, early):
            """ISA branch of scale s, followed -- for the coarser scales, which finish long before the 1/3 scale --
            by the exchange convolutions of the coming CSA stage that only need this scale's output
            (aggregation.py:346-371): they leave the CSA stage's critical path.  The coarsest row's running sum
            (go_folded) travels from stream to stream with an event."""
            def go():
                y = blocks[0](x, y1)
                for blk in blocks[1:]:
                    y = blk(y)
                if fuse is None or s == 0 or not EARLY_EXCHANGE:
                    return y
                for i, row in enumerate(fuse):
                    chain = row[s] if s < len(row) else ()
                    if folded(i, row):
                        if s == i:                                  # identity term starts the running sum
                            ev = torch.cuda.Event()
                            ev.record(torch.cuda.current_stream(dev))
                            early[("acc", i)] = (y, ev)
                        elif chain and ("acc", i) in early:
                            acc, ev = early[("acc", i)]
                            torch.cuda.current_stream(dev).wait_event(ev)
                            acc = run_chain(chain, y, residual=acc, act=ops.ACT_NONE)
                            ev = torch.cuda.Event()
                            ev.record(torch.cuda.current_stream(dev))
                            early[("acc", i)] = (acc, ev)
                            early[(i, s)] = True
                    elif chain:
                        early[(i, s)] = run_chain(chain, y)
                return y
            return go

        if nhwc:
            xs = list(cost_volume)
        else:
            xs = fork_join(dev, [(lambda c=c: ops.nchw_to_nhwc(c)) for c in cost_volume])
        pre = [None] * len(xs)                  # conv1 outputs of the next stage, where the CSA launch produced them
        for si, (branches, fuse, slope) in enumerate(self.stages):
            # ISA: the scales are independent -> one stream each
            # (coarsest scale first in program order: the folded row's running sum starts there)
            early = {}
            order = [0] + list(range(len(branches) - 1, 0, -1))
            got = fork_join(dev, [branch(s, branches[s], xs[s], pre[s], fuse, early) for s in order])
            xs = [got[order.index(s)] for s in range(len(branches))]
            pre = [None] * len(xs)
            if fuse is None:
                continue
            nxt = self.stages[si + 1][0] if si + 1 < len(self.stages) else None
            # last module + disparity requested: its sum, the final 1x1 convolution and the soft-argmin in ONE launch
            # (nothing else runs at that point, so the whole-SM CTAs of the fused kernel cost nothing)
            to_disp = (disparity and nxt is None and len(fuse) == 1 and len(self.final) == 1 and FUSE_FINAL
                       and self.disparity_fusable())
            # CSA: output scale i needs every input scale; the output scales are independent
            def fuse_row(i, row, xs=xs, nxt=nxt, to_disp=to_disp, early=early):
                def go_folded():
                    # coarsest output scale: every term has the output's size, so the sum needs no resize kernel --
                    # each exchange chain's last convolution adds the running sum as its residual and the last one
                    # applies the LeakyReLU (aggregation.py:387-400; (t0 + t1) + t2 is evaluated as t0 + (t1 + t2)).
                    # The stage's critical chain (1/3 -> 1/6 -> 1/12) runs last and loses the csa_fuse launch.
                    chains = sorted([(j, c) for j, c in enumerate(row) if c and (i, j) not in early],
                                    key=lambda jc: len(jc[1]))

                    def head(chain, x):                 # everything but the last convolution of a chain
                        for conv in chain[:-1]:
                            x = conv.single(x) if conv.tmem_eligible() else conv(x)
                        return x

                    def shorter():                      # the shorter chains, summed onto the identity term
                        acc = early[("acc", i)][0] if ("acc", i) in early else xs[i]
                        for j, chain in chains[:-1]:
                            acc = chain[-1].single(head(chain, xs[j]), residual=acc, act=ops.ACT_NONE)
                        return acc
                    jl, longest = chains[-1]
                    if len(longest) > 1 and len(chains) > 1 and FOLD_PARALLEL:
                        # the long chain's head does not need the running sum: next to the shorter chains
                        t, acc = fork_join(dev, [lambda: head(longest, xs[jl]), shorter], first=len(fuse))
                    else:
                        acc = shorter()
                        t = head(longest, xs[jl])
                    return longest[-1].single(t, residual=acc, act=ops.ACT_LEAKY, slope=slope), None
                if folded(i, row):
                    return go_folded

                def go():
                    # the last conv of every exchange chain of this row in ONE multi-problem launch (they are
                    # independent and all produce this row's channel count); longer chains run their head first
                    terms, last, where = [], [], []
                    for j, chain in enumerate(row):
                        if (i, j) in early:                         # computed inside the ISA branch of scale j
                            terms.append(early[(i, j)])
                            continue
                        t = xs[j]
                        for conv in chain[:-1]:
                            t = conv.single(t) if conv.tmem_eligible() else conv(t)
                        if chain and chain[-1].tmem_eligible():
                            t = chain[-1].single(t)                 # strided 3x3: TMEM-A kernel, own launch
                            chain = ()
                        terms.append(t)
                        if chain:
                            last.append((chain[-1], t))
                            where.append(j)
                    if len(last) == 1:
                        terms[where[0]] = last[0][0](last[0][1])
                    elif len(last) > 1 and BATCH_EXCHANGE:
                        bn = max(ops.natural_bn(c.Cout // c.groups) for c, _ in last)
                        outs = ops.conv_batch([c.problem(t, bn) for c, t in last], bn=bn)
                        for j, o in zip(where, outs):
                            terms[j] = o
                    else:
                        for j, (c, t) in zip(where, last):
                            terms[j] = c(t)
                    if to_disp and ops.csa_conv1_supported(terms, self.final[0].Cout, force=True):
                        f = self.final[0]
                        return ops.csa_conv1_nhwc(terms, slope, f.wpack, f.Cout, f.scale, f.shift,
                                                  ops.ACT_SOFTARGMIN, bias=f.bias, keep_sum=False)    # (None, disparity)
                    if nxt is not None and i < len(nxt) and nxt[i][0].conv1_fusable() and \
                            ops.csa_conv1_supported(terms, nxt[i][0].c1.Cout):
                        c1 = nxt[i][0].c1           # the next module's conv1 consumes the sum inside the same launch
                        return ops.csa_conv1_nhwc(terms, slope, c1.wpack, c1.Cout, c1.scale, c1.shift, c1.act)
                    return ops.csa_fuse_nhwc(terms, slope), None
                return go
            got = fork_join(dev, [fuse_row(i, row) for i, row in enumerate(fuse)])
            if to_disp and got[0][0] is None:
                return [got[0][1]]                  # the disparity of the finest scale
            xs, pre = [g[0] for g in got], [g[1] for g in got]
            pre += [None] * (len(branches) - len(pre))
        if disparity:
            # DisparityEstimation (similarity volume: softmax over the candidates, estimation.py:19-28) in the epilogue of
            # the final 1x1: the aggregated volume is never written
            return fork_join(dev, [(lambda s=s, conv=conv: conv(xs[s], act=ops.ACT_SOFTARGMIN))
                                   for s, conv in enumerate(self.final)])
        return fork_join(dev, [(lambda s=s, conv=conv: conv(xs[s], out_nchw=True)) for s, conv in enumerate(self.final)])


def run(agg, cost_volume, nhwc=False, disparity=False):
    """disparity=True: return soft-argmin disparities [B,H,W] (similarity volumes) instead of the aggregated volumes
    when the executor can fuse them (FusedAggregation.disparity_fusable), else None -- the caller then asks again."""
    fused = getattr(agg, "_aanet_fused", None)
    if fused is None or fused.shapes != _shape_key(agg):
        fused = FusedAggregation(agg)
        agg._aanet_fused = fused
    elif fused.key != _state_key(agg):
        # weights changed in place (or were re-assigned with the same shapes): re-pack into the SAME buffers so
        # that CUDA graphs captured earlier keep reading valid, current data
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("aanet_b200.fused: weights changed since the last eager run; run one eager forward "
                               "before capturing (the re-pack must not become part of the graph)")
        fused.refresh(agg)
    if disparity and not fused.disparity_fusable():
        return None
    return fused(cost_volume, nhwc, disparity)
