"""torch.autograd front-ends of the C-ABI kernels.

PyTorch is plumbing here (device memory, streams, autograd graph); the arithmetic is in
libaanet_b200.so.  Every function requires CUDA float32 tensors and raises NotImplementedError for
CPU tensors, mirroring nets/deform_conv/deform_conv.py:135-136,153-154 of the reference.
"""
import ctypes
import os

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib

LAUNCHES = 0   # number of C-ABI kernels-launching calls made by this process (bench bookkeeping)
FORCE_GENERIC_MDCN = False   # tests: route mdconv forward to the shape-generic FFMA kernel (ws = NULL)


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


def _prep(t, name):
    if t is None:
        return None
    if not t.is_cuda:
        raise NotImplementedError("aanet_b200.%s: CUDA tensors only (no CPU fallback)" % name)
    if t.dtype != torch.float32:
        raise TypeError("aanet_b200.%s: float32 only, got %s" % (name, t.dtype))
    return t.contiguous()


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


# ------------------------------------------------------------------------------------ correlation
class _Correlation(Function):
    @staticmethod
    def forward(ctx, left, right, max_disp):
        left, right = _prep(left, "correlation"), _prep(right, "correlation")
        if left.dim() != 4 or left.shape != right.shape:
            raise ValueError("correlation: left/right must be [B,C,H,W] of equal shape")
        B, C, H, W = left.shape
        out = left.new_empty(B, max_disp, H, W)
        with torch.cuda.device(left.device):
            _lib.check(_lib.load().aanet_corr_fwd(_ptr(left), _ptr(right), _ptr(out), B, C, H, W,
                                                  max_disp, _stream(left)), "aanet_corr_fwd")
        _count()
        ctx.save_for_backward(left, right)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        left, right = ctx.saved_tensors
        g = _prep(g, "correlation")
        B, C, H, W = left.shape
        gl, gr = torch.empty_like(left), torch.empty_like(right)
        with torch.cuda.device(left.device):
            _lib.check(_lib.load().aanet_corr_bwd(_ptr(left), _ptr(right), _ptr(g), _ptr(gl), _ptr(gr),
                                                  B, C, H, W, g.shape[1], _stream(left)), "aanet_corr_bwd")
        _count()
        return gl, gr, None


def correlation(left, right, max_disp):
    """cost[b,d,h,w] = mean_c left[b,c,h,w]*right[b,c,h,w-d], 0 where w<d (nets/cost.py:40-48).
    bfloat16 features select the bf16 cost-volume variant (inference only, fp32 volume)."""
    if left.dtype == torch.bfloat16:
        return correlation_bf16(left, right, max_disp)
    return _Correlation.apply(left, right, int(max_disp))


def correlation_nhwc(left, right, max_disp):
    """The same volume written channels-last, [B,H,W,D], for the fused aggregation executor (inference only;
    needs max_disp % 4 == 0 and max_disp <= 128 -- AanetError otherwise)."""
    left, right = _prep(left, "correlation_nhwc"), _prep(right, "correlation_nhwc")
    if left.shape != right.shape or left.dim() != 4:
        raise ValueError("correlation_nhwc: left/right must be [B,C,H,W] tensors of equal shape")
    B, C, H, W = left.shape
    out = left.new_empty(B, H, W, int(max_disp))
    with torch.cuda.device(left.device):
        _lib.check(_lib.load().aanet_corr_fwd_nhwc(_ptr(left), _ptr(right), _ptr(out), B, C, H, W, int(max_disp),
                                                   _stream(left)), "aanet_corr_fwd_nhwc")
    _count()
    return out


def correlation_bf16(left, right, max_disp):
    """BASELINE config 5 variant: bf16 features, fp32 accumulation and fp32 volume (no autograd)."""
    if not left.is_cuda:
        raise NotImplementedError("aanet_b200.correlation_bf16: CUDA tensors only (no CPU fallback)")
    if left.dtype != torch.bfloat16 or right.dtype != torch.bfloat16 or left.shape != right.shape:
        raise TypeError("correlation_bf16: left/right must be bfloat16 tensors of equal shape")
    left, right = left.contiguous(), right.contiguous()
    B, C, H, W = left.shape
    out = torch.empty(B, int(max_disp), H, W, dtype=torch.float32, device=left.device)
    with torch.cuda.device(left.device):
        _lib.check(_lib.load().aanet_corr_fwd_bf16(_ptr(left), _ptr(right), _ptr(out), B, C, H, W, int(max_disp),
                                                   _stream(left)), "aanet_corr_fwd_bf16")
    _count()
    return out


class _Cost5d(Function):
    @staticmethod
    def forward(ctx, left, right, max_disp, mode):
        left, right = _prep(left, "cost_volume_5d"), _prep(right, "cost_volume_5d")
        if left.dim() != 4 or left.shape != right.shape:
            raise ValueError("cost_volume_5d: left/right must be [B,C,H,W] of equal shape")
        B, C, H, W = left.shape
        out = left.new_empty(B, C * (2 if mode else 1), max_disp, H, W)
        with torch.cuda.device(left.device):
            _lib.check(_lib.load().aanet_cost5d_fwd(_ptr(left), _ptr(right), _ptr(out), B, C, H, W, max_disp, mode,
                                                    _stream(left)), "aanet_cost5d_fwd")
        _count()
        ctx.meta = (B, C, H, W, max_disp, mode)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        B, C, H, W, D, mode = ctx.meta
        g = _prep(g, "cost_volume_5d")
        gl, gr = g.new_empty(B, C, H, W), g.new_empty(B, C, H, W)
        with torch.cuda.device(g.device):
            _lib.check(_lib.load().aanet_cost5d_bwd(_ptr(g), _ptr(gl), _ptr(gr), B, C, H, W, D, mode, _stream(g)),
                       "aanet_cost5d_bwd")
        _count()
        return gl, gr, None, None


def cost_volume_5d(left, right, max_disp, kind):
    """'difference' [B,C,D,H,W] or 'concat' [B,2C,D,H,W] volume (nets/cost.py:22-38), zero where w < d."""
    if kind not in ("difference", "concat"):
        raise NotImplementedError(kind)
    return _Cost5d.apply(left, right, int(max_disp), 1 if kind == "concat" else 0)


# ------------------------------------------------------------------------------------ soft-argmin
class _SoftArgmin(Function):
    @staticmethod
    def forward(ctx, cost, similarity):
        cost = _prep(cost, "soft_argmin")
        B, D, H, W = cost.shape
        disp = cost.new_empty(B, H, W)
        with torch.cuda.device(cost.device):
            _lib.check(_lib.load().aanet_softargmin_fwd(_ptr(cost), _ptr(disp), B, D, H, W,
                                                        int(similarity), _stream(cost)),
                       "aanet_softargmin_fwd")
        _count()
        ctx.save_for_backward(cost)
        ctx.similarity = int(similarity)
        return disp

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        cost, = ctx.saved_tensors
        g = _prep(g, "soft_argmin")
        B, D, H, W = cost.shape
        gc = torch.empty_like(cost)
        with torch.cuda.device(cost.device):
            _lib.check(_lib.load().aanet_softargmin_bwd(_ptr(cost), _ptr(g), _ptr(gc), B, D, H, W,
                                                        ctx.similarity, _stream(cost)),
                       "aanet_softargmin_bwd")
        _count()
        return gc, None


def soft_argmin(cost, similarity=True):
    """disp = sum_d d*softmax_d(+-cost) (nets/estimation.py:13-30)."""
    return _SoftArgmin.apply(cost, bool(similarity))


def refine_frontend(low_disp, left_img, right_img):
    """(low_disp [B,h,w], left [B,C,H,W], right [B,C,H,W]) -> (cat(warped_right - left, left) [B,2C,H,W],
    disp [B,1,H,W]): upsample + rescale + disp_warp + error + concat of refinement.py:80-95 / warp.py:41-64 in
    one launch and without the reference's host synchronisation.  Inference only (no autograd)."""
    low = _prep(low_disp, "refine_frontend")
    left, right = _prep(left_img, "refine_frontend"), _prep(right_img, "refine_frontend")
    if low.dim() != 3 or left.dim() != 4 or left.shape != right.shape or low.shape[0] != left.shape[0]:
        raise ValueError("refine_frontend: low_disp [B,h,w], left/right [B,C,H,W] expected")
    B, C, H, W = left.shape
    concat = left.new_empty(B, 2 * C, H, W)
    disp = left.new_empty(B, 1, H, W)
    with torch.cuda.device(left.device):
        _lib.check(_lib.load().aanet_refine_frontend_fwd(_ptr(low), _ptr(left), _ptr(right), _ptr(concat), _ptr(disp),
                                                         B, C, low.shape[1], low.shape[2], H, W, _stream(left)),
                   "aanet_refine_frontend_fwd")
    _count()
    return concat, disp


# ------------------------------------------------------------------------------------ mdconv
def _out_hw(H, W, kh, kw, stride, pad, dil):
    return ((H + 2 * pad - (dil * (kh - 1) + 1)) // stride + 1,
            (W + 2 * pad - (dil * (kw - 1) + 1)) // stride + 1)


def _mdcn_forward(x, offset, mask, weight, bias, stride, pad, dil, groups, dg,
                  post_scale=None, post_shift=None, relu=False):
    B, Cin, H, W = x.shape
    Cout, cg, kh, kw = weight.shape
    if Cin != cg * groups:
        raise _lib.AanetError("Input shape and kernel channels wont match: (%d vs %d)." % (Cin, cg * groups))
    Ho, Wo = _out_hw(H, W, kh, kw, stride, pad, dil)
    if Ho <= 0 or Wo <= 0:
        raise ValueError("convolution input is too small (output would be %dx%d)" % (Ho, Wo))
    if tuple(offset.shape) != (B, dg * 2 * kh * kw, Ho, Wo):
        raise _lib.AanetError("offset must be [%d,%d,%d,%d], got %s" % (B, dg * 2 * kh * kw, Ho, Wo,
                                                                         tuple(offset.shape)))
    if mask is not None and tuple(mask.shape) != (B, dg * kh * kw, Ho, Wo):
        raise _lib.AanetError("mask must be [%d,%d,%d,%d], got %s" % (B, dg * kh * kw, Ho, Wo,
                                                                       tuple(mask.shape)))
    out = x.new_empty(B, Cout, Ho, Wo)
    lib = _lib.load()
    nbytes = 0 if FORCE_GENERIC_MDCN else lib.aanet_mdcn_workspace_bytes(
        0, B, Cin, H, W, Cout, kh, kw, stride, pad, dil, groups, dg)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device) if nbytes else None
    with torch.cuda.device(x.device):
        _lib.check(lib.aanet_mdcn_fwd(
            _ptr(x), _ptr(offset), _ptr(mask), _ptr(weight), _ptr(bias), _ptr(out),
            B, Cin, H, W, Cout, kh, kw, stride, pad, dil, groups, dg,
            _ptr(post_scale), _ptr(post_shift), int(relu), _ptr(ws), nbytes, _stream(x)), "aanet_mdcn_fwd")
    _count(3 if nbytes else 1)
    return out


def _mdcn_backward(x, offset, mask, weight, with_bias, g, stride, pad, dil, groups, dg):
    B, Cin, H, W = x.shape
    Cout, _, kh, kw = weight.shape
    gx, goff = torch.empty_like(x), torch.empty_like(offset)
    gmask = torch.empty_like(mask) if mask is not None else None
    gw = torch.empty_like(weight)
    gb = weight.new_empty(Cout) if with_bias else None
    lib = _lib.load()
    nbytes = lib.aanet_mdcn_workspace_bytes(1, B, Cin, H, W, Cout, kh, kw, stride, pad, dil, groups, dg)
    ws = torch.empty(max(nbytes, 4), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.aanet_mdcn_bwd(
            _ptr(x), _ptr(offset), _ptr(mask), _ptr(weight), _ptr(g), _ptr(gx), _ptr(goff), _ptr(gmask),
            _ptr(gw), _ptr(gb), B, Cin, H, W, Cout, kh, kw, stride, pad, dil, groups, dg,
            _ptr(ws), nbytes, _stream(x)), "aanet_mdcn_bwd")
    _count(4)
    return gx, goff, gmask, gw, gb


def _single_int(v, name):
    if isinstance(v, (tuple, list)):
        if len(set(v)) != 1:
            raise _lib.AanetError("%s must be the same in h and w, got %s" % (name, (v,)))
        v = v[0]
    return int(v)


class ModulatedDeformConvFunction(Function):
    """Same call signature as the reference's (nets/deform_conv/deform_conv.py:113-171)."""

    @staticmethod
    def forward(ctx, input, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
                deformable_groups=1):
        ctx.cfg = (_single_int(stride, "stride"), _single_int(padding, "padding"),
                   _single_int(dilation, "dilation"), int(groups), int(deformable_groups))
        ctx.with_bias = bias is not None
        input, offset, mask = _prep(input, "mdconv"), _prep(offset, "mdconv"), _prep(mask, "mdconv")
        weight, bias = _prep(weight, "mdconv"), _prep(bias, "mdconv")
        ctx.save_for_backward(input, offset, mask, weight)
        return _mdcn_forward(input, offset, mask, weight, bias, *ctx.cfg)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        input, offset, mask, weight = ctx.saved_tensors
        g = _prep(grad_output, "mdconv")
        gx, goff, gmask, gw, gb = _mdcn_backward(input, offset, mask, weight, ctx.with_bias, g, *ctx.cfg)
        return gx, goff, gmask, gw, gb, None, None, None, None, None


class DeformConvFunction(Function):
    """DCNv1 (nets/deform_conv/deform_conv.py:12-110): the same kernels with mask == 1.  The
    reference's im2col_step batching constraint (B % step == 0, :47-49) does not exist here; the
    argument is accepted and ignored."""

    @staticmethod
    def forward(ctx, input, offset, weight, stride=1, padding=0, dilation=1, groups=1,
                deformable_groups=1, im2col_step=64):
        if input is not None and input.dim() != 4:
            raise ValueError("Expected 4D tensor as input, got {}D tensor instead.".format(input.dim()))
        ctx.cfg = (_single_int(stride, "stride"), _single_int(padding, "padding"),
                   _single_int(dilation, "dilation"), int(groups), int(deformable_groups))
        input, offset, weight = _prep(input, "deform_conv"), _prep(offset, "deform_conv"), \
            _prep(weight, "deform_conv")
        ctx.save_for_backward(input, offset, weight)
        return _mdcn_forward(input, offset, None, weight, None, *ctx.cfg)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        input, offset, weight = ctx.saved_tensors
        g = _prep(grad_output, "deform_conv")
        gx, goff, _, gw, _ = _mdcn_backward(input, offset, None, weight, False, g, *ctx.cfg)
        return gx, goff, gw, None, None, None, None, None, None


modulated_deform_conv = ModulatedDeformConvFunction.apply
deform_conv = DeformConvFunction.apply


def modulated_deform_conv_fused(x, offset, mask, weight, bias, stride, padding, dilation, groups,
                                deformable_groups, post_scale, post_shift, relu):
    """Inference-only: mdconv with a per-channel affine (folded BatchNorm) + ReLU epilogue."""
    x, offset, mask = _prep(x, "mdconv"), _prep(offset, "mdconv"), _prep(mask, "mdconv")
    return _mdcn_forward(x, offset, mask, _prep(weight, "mdconv"), _prep(bias, "mdconv"), stride, padding,
                         dilation, groups, deformable_groups, _prep(post_scale, "mdconv"),
                         _prep(post_shift, "mdconv"), relu)


# ------------------------------------------------------------------------------------ dense conv
ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2


def conv2d_fused(x, weight, bias=None, scale=None, shift=None, residual=None, act=ACT_NONE, slope=0.2,
                 stride=1, padding=0, dilation=1, groups=1):
    """Inference-only dense convolution on the tcgen05 engine:
    act((conv(x, w) + bias) * scale + shift + residual).  Replaces cuDNN conv + BatchNorm(eval) +
    activation (+ residual add) of the ISA block / CSA exchange paths with one kernel."""
    x, weight = _prep(x, "conv2d"), _prep(weight, "conv2d")
    bias, scale, shift, residual = (_prep(t, "conv2d") for t in (bias, scale, shift, residual))
    B, Cin, H, W = x.shape
    Cout, cg, kh, kw = weight.shape
    if Cin != cg * groups:
        raise _lib.AanetError("Input shape and kernel channels wont match: (%d vs %d)." % (Cin, cg * groups))
    Ho, Wo = _out_hw(H, W, kh, kw, stride, padding, dilation)
    out = x.new_empty(B, Cout, Ho, Wo)
    if residual is not None and residual.shape != out.shape:
        raise ValueError("conv2d_fused: residual must have the output's shape")
    lib = _lib.load()
    nbytes = lib.aanet_conv2d_workspace_bytes(B, Cin, H, W, Cout, kh, kw, stride, padding, dilation, groups)
    if nbytes == 0:
        raise _lib.AanetError("conv2d_fused needs Cin %% 4 == 0 and Cin/groups %% 4 == 0 (got Cin=%d, groups=%d)"
                              % (Cin, groups))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.aanet_conv2d_fwd(_ptr(x), _ptr(weight), _ptr(bias), _ptr(scale), _ptr(shift),
                                        _ptr(residual), int(act), float(slope), _ptr(out), B, Cin, H, W, Cout,
                                        kh, kw, stride, padding, dilation, groups, _ptr(ws), nbytes,
                                        _stream(x)), "aanet_conv2d_fwd")
    _count(3)
    return out


# ------------------------------------------------------------------------------------ channels-last engine
ACT_OFFSET_MASK = 3
ACT_SOFTARGMIN = 4      # conv_batch: the epilogue reduces the output channels to the soft-argmin disparity, out [B,Ho,Wo]


def nchw_to_nhwc(x):
    """[B,C,H,W] -> [B,H,W,C] (contiguous) with the library's tiled transpose."""
    x = _prep(x, "nchw_to_nhwc")
    B, C, H, W = x.shape
    out = x.new_empty(B, H, W, C)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().aanet_nchw_to_nhwc(_ptr(x), _ptr(out), B, C, H * W, _stream(x)), "aanet_nchw_to_nhwc")
    _count()
    return out


def nhwc_to_nchw(x):
    x = _prep(x, "nhwc_to_nchw")
    B, H, W, C = x.shape
    out = x.new_empty(B, C, H, W)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().aanet_nhwc_to_nchw(_ptr(x), _ptr(out), B, C, H * W, _stream(x)), "aanet_nhwc_to_nchw")
    _count()
    return out


def natural_bn(out_per_group):
    """N-tile width the engine picks for a layer on its own: multiple of 16, <= 64, wider layers split evenly."""
    n_tiles = (out_per_group + 63) // 64
    per = (out_per_group + n_tiles - 1) // n_tiles
    return (per + 15) // 16 * 16


def pack_conv_weight(weight, groups=1, bn=0, out=None):
    """tf32 hi/lo split + 128-byte swizzle of a [Cout, Cin/groups, kh, kw] weight for the tcgen05 engine.
    bn = N-tile width of the launch the layer will run in (0: the layer's own natural width).
    out: an earlier result for the same shape, overwritten in place (its address may be baked into CUDA graphs)."""
    weight = _prep(weight.detach(), "pack_conv_weight")
    Cout, cg, kh, kw = weight.shape
    lib = _lib.load()
    n = lib.aanet_conv_wpack_bytes(Cout, cg * groups, kh, kw, groups, bn)
    if n == 0:
        raise _lib.AanetError("conv engine needs Cin/groups %% 4 == 0 (got Cin=%d, groups=%d)" % (cg * groups, groups))
    if out is not None and (out.numel() != n or out.device != weight.device or out.dtype != torch.uint8):
        raise ValueError("pack_conv_weight: `out` does not match this weight's packed size")
    wpack = torch.empty(n, dtype=torch.uint8, device=weight.device) if out is None else out
    with torch.cuda.device(weight.device):
        _lib.check(lib.aanet_conv_pack_weights(_ptr(weight), _ptr(wpack), Cout, cg * groups, kh, kw, groups, bn,
                                               _stream(weight)), "aanet_conv_pack_weights")
    _count()
    return wpack


def _dp(t):
    return None if t is None else t.data_ptr()


def _cl(t, name, dev, shape=None):
    """Descriptor operand of the channels-last engine: CUDA fp32, contiguous, on `dev`, optionally of `shape`.
    The engine sees raw pointers, so anything else would be silent garbage or an illegal address."""
    if t is None:
        return None
    if not t.is_cuda:
        raise NotImplementedError("aanet_b200.conv_batch: %s must be a CUDA tensor (no CPU fallback)" % name)
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise TypeError("aanet_b200.conv_batch: %s must be contiguous float32, got %s%s"
                        % (name, t.dtype, "" if t.is_contiguous() else " (non-contiguous)"))
    if t.device != dev:
        raise ValueError("aanet_b200.conv_batch: %s is on %s, x is on %s" % (name, t.device, dev))
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError("aanet_b200.conv_batch: %s must be %s, got %s" % (name, tuple(shape), tuple(t.shape)))
    return t


def _fill_desc(d, q, out):
    """Fill one aanet_conv_desc from a problem dict (see conv_batch); `out` is the output tensor or None."""
    x = q["x"]
    if x.dim() != 4:
        raise ValueError("aanet_b200.conv_batch: x must be channels-last [B,H,W,C]")
    dev = x.device
    _cl(x, "x", dev)
    B, H, W, Cin = x.shape
    Ho, Wo = _out_hw(H, W, q["kh"], q["kw"], q["stride"], q["pad"], q["dil"])
    tail = q.get("tail")
    c_out = tail["Cout"] if tail else q["Cout"]
    oshape = (B, c_out, Ho, Wo) if q.get("out_nchw") else (B, Ho, Wo, c_out)
    if int(q.get("act", ACT_NONE)) == ACT_SOFTARGMIN:
        oshape = (B, Ho, Wo)
    om = q.get("offmask")
    for nm in ("bias", "scale", "shift"):
        _cl(q.get(nm), nm, dev, (q["Cout"],))
    _cl(q.get("residual"), "residual", dev, oshape)
    if om is not None:
        if om.dim() != 4 or (tuple(om.shape[2:]) if q.get("om_nchw") else tuple(om.shape[1:3])) != (Ho, Wo) \
                or om.shape[0] != B:
            raise ValueError("aanet_b200.conv_batch: offmask must be [B,om,Ho,Wo] (om_nchw) or [B,Ho,Wo,om]")
        _cl(om, "offmask", dev)
    if q["wpack"].device != dev:
        raise ValueError("aanet_b200.conv_batch: packed weights are on %s, x is on %s" % (q["wpack"].device, dev))
    d.x, d.wpack = x.data_ptr(), q["wpack"].data_ptr()
    d.out = None if out is None else out.data_ptr()
    d.bias, d.scale, d.shift = _dp(q.get("bias")), _dp(q.get("scale")), _dp(q.get("shift"))
    d.residual, d.offmask = _dp(q.get("residual")), _dp(om)
    d.om_nchw = int(bool(q.get("om_nchw")))
    d.om_channels = 0 if om is None else (om.shape[1] if d.om_nchw else om.shape[-1])
    d.B, d.Cin, d.H, d.W, d.Cout, d.kh, d.kw = B, Cin, H, W, q["Cout"], q["kh"], q["kw"]
    d.stride, d.pad, d.dil, d.groups, d.dg = q["stride"], q["pad"], q["dil"], q.get("groups", 1), q.get("dg", 1)
    d.act, d.slope = int(q.get("act", ACT_NONE)), float(q.get("slope", 0.2))
    d.n_offset_ch, d.mask_scale = int(q.get("n_offset_ch", 0)), float(q.get("mask_scale", 1.0))
    d.out_nchw = int(bool(q.get("out_nchw")))
    if tail:
        if q.get("residual") is not None or q.get("out_nchw"):
            raise ValueError("aanet_b200.conv_batch: a problem with a fused tail has no main residual / NCHW output")
        for nm in ("scale", "shift"):
            _cl(tail.get(nm), "tail " + nm, dev, (tail["Cout"],))
        _cl(tail.get("residual"), "tail residual", dev, oshape)
        d.tail_wpack = tail["wpack"].data_ptr()
        d.tail_scale, d.tail_shift = _dp(tail.get("scale")), _dp(tail.get("shift"))
        d.tail_residual = _dp(tail.get("residual"))
        d.tail_cout, d.tail_act = int(tail["Cout"]), int(tail.get("act", ACT_NONE))
    return oshape


def conv_tail_supported(problem, deform=False):
    """Can `problem` (a conv_batch dict with a "tail" entry: the bottleneck's trailing 1x1 + BN + residual + act) run
    as ONE launch of the tensor-memory kernels?"""
    d = _lib.ConvDesc()
    _fill_desc(d, problem, problem["x"])          # any valid pointer: the query does not launch
    return bool(_lib.load().aanet_conv_tail_supported(ctypes.byref(d), int(deform)))


def conv_batch(problems, deform=False, bn=0):
    """Run 1..3 problems (dicts: x, wpack, Cout, kh, kw, stride, pad, dil [, bias, scale, shift, residual, act,
    slope, groups, dg, offmask, om_nchw, out_nchw, n_offset_ch, mask_scale, tail]) as one persistent engine launch;
    returns their outputs."""
    n = len(problems)
    descs = (_lib.ConvDesc * n)()
    outs = []
    for d, q in zip(descs, problems):
        oshape = _fill_desc(d, q, None)
        out = q["x"].new_empty(oshape)
        d.out = out.data_ptr()
        outs.append(out)
    x0 = problems[0]["x"]
    with torch.cuda.device(x0.device):
        _lib.check(_lib.load().aanet_conv_batch_nhwc(ctypes.cast(descs, ctypes.c_void_p), n, int(deform), int(bn),
                                                     _stream(x0)), "aanet_conv_batch_nhwc")
    _count()
    return outs


def conv2d_nhwc(x, wpack, Cout, kh, kw, bias=None, scale=None, shift=None, residual=None, act=ACT_NONE,
                slope=0.2, stride=1, padding=0, dilation=1, groups=1, out_nchw=False, n_offset_ch=0,
                mask_scale=1.0, tail=None):
    """Dense convolution on channels-last activations x [B,H,W,Cin] with pre-packed weights."""
    return conv_batch([dict(x=x, wpack=wpack, Cout=Cout, kh=kh, kw=kw, bias=bias, scale=scale, shift=shift,
                            residual=residual, act=act, slope=slope, stride=stride, pad=padding, dil=dilation,
                            groups=groups, out_nchw=out_nchw, n_offset_ch=n_offset_ch, mask_scale=mask_scale,
                            tail=tail)])[0]


def mdcn_nhwc(x, offmask, wpack, Cout, kh, kw, bias=None, scale=None, shift=None, relu=False, stride=1,
              padding=0, dilation=1, groups=1, deformable_groups=1, out_nchw=False, om_nchw=False, tail=None):
    """DCNv2 on channels-last x [B,H,W,Cin] with offsets+mask in one tensor: channels-last [B,Ho,Wo,om] or,
    with om_nchw, channel planes [B,om,Ho,Wo] (one L1 wavefront per offset load instead of ~20)."""
    return conv_batch([dict(x=x, offmask=offmask, wpack=wpack, Cout=Cout, kh=kh, kw=kw, bias=bias, scale=scale,
                            shift=shift, act=ACT_RELU if relu else ACT_NONE, stride=stride, pad=padding,
                            dil=dilation, groups=groups, dg=deformable_groups, out_nchw=out_nchw,
                            om_nchw=om_nchw, tail=tail)], deform=True)[0]


def csa_fuse_nhwc(terms, slope=0.2):
    """Channels-last csa_fuse (inference): terms [B,h,w,C] -> [B,H,W,C] of terms[0]'s size."""
    B, H, W, C = terms[0].shape
    n = len(terms)
    for t in terms:
        _cl(t, "csa term", terms[0].device)
        if t.dim() != 4 or t.shape[0] != B or t.shape[3] != C:
            raise ValueError("csa_fuse_nhwc: all terms must be [B,h,w,C] with the same B and C")
    th = (ctypes.c_int * n)(*[t.shape[1] for t in terms])
    tw = (ctypes.c_int * n)(*[t.shape[2] for t in terms])
    out = terms[0].new_empty(B, H, W, C)
    with torch.cuda.device(out.device):
        _lib.check(_lib.load().aanet_csa_fuse_nhwc(_term_arrays(terms), th, tw, n, _ptr(out), B, C, H, W,
                                                   float(slope), _stream(out)), "aanet_csa_fuse_nhwc")
    _count()
    return out


def csa_conv1_supported(terms, Cout, force=False):
    """Can csa_conv1_nhwc run these terms (first one sets the output size) and a 1x1 convolution to Cout channels?
    Inside the aggregation stages the fused launch is opt-in (AANET_CSA_CONV1=1: it measured slower there);
    force=True asks for the shape check only (the last module, where nothing runs next to it)."""
    if (not force and os.environ.get("AANET_CSA_CONV1", "0") != "1") or not 1 <= len(terms) <= 3 or Cout not in (32, 64):
        return False
    B, H, W, C = terms[0].shape
    if C % 32:
        return False
    return all((t.shape[1] == H and t.shape[2] == W) or (t.shape[1] < H and t.shape[2] < W) for t in terms)


def csa_conv1_nhwc(terms, slope, wpack, Cout, scale=None, shift=None, act=ACT_RELU, bias=None, keep_sum=True):
    """CSA resize-and-sum + LeakyReLU (aggregation.py:387-400) and the following 1x1 convolution + folded BN +
    activation (conv1 of the next bottleneck, deform.py:164-170) as one launch.  Returns (sum [B,H,W,C] or None when
    keep_sum is False, conv output [B,H,W,Cout] -- or the soft-argmin disparity [B,H,W] for act = ACT_SOFTARGMIN)."""
    B, H, W, C = terms[0].shape
    n = len(terms)
    dev = terms[0].device
    for t in terms:
        _cl(t, "csa term", dev)
        if t.dim() != 4 or t.shape[0] != B or t.shape[3] != C:
            raise ValueError("csa_conv1_nhwc: all terms must be [B,h,w,C] with the same B and C")
    for nm, v in (("bias", bias), ("scale", scale), ("shift", shift)):
        _cl(v, nm, dev, (Cout,))
    th = (ctypes.c_int * n)(*[t.shape[1] for t in terms])
    tw = (ctypes.c_int * n)(*[t.shape[2] for t in terms])
    fused = terms[0].new_empty(B, H, W, C) if keep_sum else None
    out = terms[0].new_empty((B, H, W) if int(act) == ACT_SOFTARGMIN else (B, H, W, Cout))
    with torch.cuda.device(dev):
        _lib.check(_lib.load().aanet_csa_conv1_nhwc(_term_arrays(terms), th, tw, n, float(slope), _dp(fused),
                                                    wpack.data_ptr(), _dp(bias), _dp(scale), _dp(shift), int(act),
                                                    _ptr(out), B, C, Cout, H, W, _stream(out)), "aanet_csa_conv1_nhwc")
    _count()
    return fused, out


# ------------------------------------------------------------------------------------ CSA fuse
def _term_arrays(tensors):
    n = len(tensors)
    ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() if t is not None else None for t in tensors])
    return ptrs


class _CsaFuse(Function):
    @staticmethod
    def forward(ctx, slope, *terms):
        terms = [_prep(t, "csa_fuse") for t in terms]
        B, C, H, W = terms[0].shape
        for t in terms:
            if t.shape[:2] != terms[0].shape[:2]:
                raise ValueError("csa_fuse: all terms must share [B,C]")
        n = len(terms)
        th = (ctypes.c_int * n)(*[t.shape[2] for t in terms])
        tw = (ctypes.c_int * n)(*[t.shape[3] for t in terms])
        out = terms[0].new_empty(B, C, H, W)
        with torch.cuda.device(out.device):
            _lib.check(_lib.load().aanet_csa_fuse_fwd(_term_arrays(terms), th, tw, n, _ptr(out), B, C, H, W,
                                                      float(slope), _stream(out)), "aanet_csa_fuse_fwd")
        _count()
        ctx.save_for_backward(out)
        ctx.meta = (slope, [tuple(t.shape) for t in terms])
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        out, = ctx.saved_tensors
        slope, shapes = ctx.meta
        g = _prep(g, "csa_fuse")
        B, C, H, W = out.shape
        n = len(shapes)
        grads = [out.new_empty(s) if ctx.needs_input_grad[i + 1] else None for i, s in enumerate(shapes)]
        th = (ctypes.c_int * n)(*[s[2] for s in shapes])
        tw = (ctypes.c_int * n)(*[s[3] for s in shapes])
        with torch.cuda.device(out.device):
            _lib.check(_lib.load().aanet_csa_fuse_bwd(_ptr(out), _ptr(g), _term_arrays(grads), th, tw, n,
                                                      B, C, H, W, float(slope), _stream(out)),
                       "aanet_csa_fuse_bwd")
        _count(n)
        return (None,) + tuple(grads)


def csa_fuse(terms, slope=0.2):
    """LeakyReLU(sum_j resize(terms[j]) to terms[0]'s size), reference order
    (nets/aggregation.py:387-400)."""
    return _CsaFuse.apply(float(slope), *terms)
