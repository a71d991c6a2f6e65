"""TEST INFRASTRUCTURE ONLY -- numpy front-end of the C oracle (oracle/aanet_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package (aanet_b200/) never does.

Each wrapper takes/returns numpy arrays (float32 or float64, C-contiguous NCHW) and calls
the matching `orc_*_f32` / `orc_*_f64` symbol.  Reference citations live next to the C code.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = [os.path.join(_HERE, "aanet_oracle.c"), os.path.join(_HERE, "aanet_oracle_impl.h")]
_LIB = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force=False):
    """gcc the C oracle into oracle/liboracle.so (no-op when up to date)."""
    if (not force and os.path.exists(_LIB)
            and all(os.path.getmtime(_LIB) >= os.path.getmtime(s) for s in _SRC)):
        return _LIB
    cmd = ["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-o", _LIB, _SRC[0], "-lm"]
    subprocess.run(cmd, check=True, cwd=_HERE)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _sfx(a):
    if a.dtype == np.float32:
        return "_f32", ctypes.c_float
    if a.dtype == np.float64:
        return "_f64", ctypes.c_double
    raise TypeError("oracle takes float32 or float64, got %s" % a.dtype)


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dt=None):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


def mdcn_out_hw(H, W, k, stride, pad, dil):
    return ((H + 2 * pad - (dil * (k - 1) + 1)) // stride + 1,
            (W + 2 * pad - (dil * (k - 1) + 1)) // stride + 1)


# ---------------------------------------------------------------- correlation
def corr_fwd(L, R, D):
    L = _c(L); R = _c(R, L.dtype)
    B, C, H, W = L.shape
    out = np.empty((B, D, H, W), L.dtype)
    s, _ = _sfx(L)
    getattr(lib(), "orc_corr_fwd" + s)(_p(L), _p(R), _p(out), B, C, H, W, D)
    return out


def corr_bwd(L, R, g):
    L = _c(L); R = _c(R, L.dtype); g = _c(g, L.dtype)
    B, C, H, W = L.shape
    D = g.shape[1]
    gL = np.empty_like(L); gR = np.empty_like(R)
    s, _ = _sfx(L)
    getattr(lib(), "orc_corr_bwd" + s)(_p(L), _p(R), _p(g), _p(gL), _p(gR), B, C, H, W, D)
    return gL, gR


def corr_pyramid(Ls, Rs, D0):
    """nets/cost.py:64-76: scale s uses max_disp // 2**s."""
    return [corr_fwd(l, r, D0 // (2 ** s)) for s, (l, r) in enumerate(zip(Ls, Rs))]


def cost5d_fwd(L, R, D, kind):
    """nets/cost.py:22-38 restated in numpy: 'difference' [B,C,D,H,W] / 'concat' [B,2C,D,H,W], zero for w < d."""
    L = np.asarray(L); R = np.asarray(R)
    B, C, H, W = L.shape
    out = np.zeros((B, C if kind == "difference" else 2 * C, D, H, W), L.dtype)
    for d in range(min(D, W)):
        if kind == "difference":
            out[:, :, d, :, d:] = L[..., d:] - R[..., :W - d]
        else:
            out[:, :C, d, :, d:] = L[..., d:]
            out[:, C:, d, :, d:] = R[..., :W - d]
    return out


def cost5d_bwd(g, kind):
    """Autograd of cost.py:22-38: (gL, gR) [B,C,H,W] from g of the forward's shape."""
    g = np.asarray(g, np.float64)
    B, Co, D, H, W = g.shape
    C = Co if kind == "difference" else Co // 2
    gL = np.zeros((B, C, H, W)); gR = np.zeros((B, C, H, W))
    for d in range(min(D, W)):
        gL[..., d:] += g[:, :C, d, :, d:]
        if kind == "difference":
            gR[..., :W - d] -= g[:, :C, d, :, d:]
        else:
            gR[..., :W - d] += g[:, C:, d, :, d:]
    return gL, gR


# ---------------------------------------------------------------- soft-argmin
def softargmin_fwd(cost, similarity=True):
    cost = _c(cost)
    B, D, H, W = cost.shape
    disp = np.empty((B, H, W), cost.dtype)
    s, _ = _sfx(cost)
    getattr(lib(), "orc_softargmin_fwd" + s)(_p(cost), _p(disp), B, D, H, W, int(bool(similarity)))
    return disp


def softargmin_bwd(cost, gdisp, similarity=True):
    cost = _c(cost); gdisp = _c(gdisp, cost.dtype)
    B, D, H, W = cost.shape
    gcost = np.empty_like(cost)
    s, _ = _sfx(cost)
    getattr(lib(), "orc_softargmin_bwd" + s)(_p(cost), _p(gdisp), _p(gcost), B, D, H, W,
                                            int(bool(similarity)))
    return gcost


# ---------------------------------------------------------------- mdconv
def mdcn_fwd(x, offset, mask, weight, bias=None, stride=1, pad=0, dil=1, groups=1, dg=1):
    x = _c(x); dt = x.dtype
    offset = _c(offset, dt); mask = _c(mask, dt); weight = _c(weight, dt); bias = _c(bias, dt)
    B, Cin, H, W = x.shape
    Cout, _, kh, kw = weight.shape
    Ho, Wo = mdcn_out_hw(H, W, kh, stride, pad, dil)
    out = np.empty((B, Cout, Ho, Wo), dt)
    s, _ = _sfx(x)
    rc = getattr(lib(), "orc_mdcn_fwd" + s)(_p(x), _p(offset), _p(mask), _p(weight), _p(bias),
                                           _p(out), B, Cin, H, W, Cout, kh, kw, stride, pad, dil,
                                           groups, dg)
    if rc:
        raise ValueError("orc_mdcn_fwd: bad arguments (%d)" % rc)
    return out


def mdcn_bwd(x, offset, mask, weight, gout, with_bias=False, stride=1, pad=0, dil=1, groups=1,
             dg=1):
    x = _c(x); dt = x.dtype
    offset = _c(offset, dt); mask = _c(mask, dt); weight = _c(weight, dt); gout = _c(gout, dt)
    B, Cin, H, W = x.shape
    Cout, _, kh, kw = weight.shape
    gx = np.empty_like(x); goff = np.empty_like(offset)
    gmask = None if mask is None else np.empty_like(mask)
    gw = np.empty_like(weight)
    gb = np.empty((Cout,), dt) if with_bias else None
    s, _ = _sfx(x)
    rc = getattr(lib(), "orc_mdcn_bwd" + s)(_p(x), _p(offset), _p(mask), _p(weight), _p(gout),
                                           _p(gx), _p(goff), _p(gmask), _p(gw), _p(gb),
                                           B, Cin, H, W, Cout, kh, kw, stride, pad, dil, groups, dg)
    if rc:
        raise ValueError("orc_mdcn_bwd: bad arguments (%d)" % rc)
    return gx, goff, gmask, gw, gb


# ---------------------------------------------------------------- CSA fuse
def _ptr_array(arrs):
    return (ctypes.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


def csa_fuse_fwd(terms, out_hw, slope=0.2):
    terms = [_c(t) for t in terms]
    dt = terms[0].dtype
    terms = [_c(t, dt) for t in terms]
    B, C = terms[0].shape[:2]
    H, W = out_hw
    th = (ctypes.c_int * len(terms))(*[t.shape[2] for t in terms])
    tw = (ctypes.c_int * len(terms))(*[t.shape[3] for t in terms])
    out = np.empty((B, C, H, W), dt)
    s, ct = _sfx(terms[0])
    getattr(lib(), "orc_csa_fuse_fwd" + s)(_ptr_array(terms), th, tw, len(terms), _p(out),
                                           B, C, H, W, ct(slope))
    return out


def csa_fuse_bwd(out, gout, term_hws, slope=0.2):
    out = _c(out); gout = _c(gout, out.dtype)
    B, C, H, W = out.shape
    gts = [np.empty((B, C, h, w), out.dtype) for (h, w) in term_hws]
    th = (ctypes.c_int * len(gts))(*[h for h, _ in term_hws])
    tw = (ctypes.c_int * len(gts))(*[w for _, w in term_hws])
    s, ct = _sfx(out)
    getattr(lib(), "orc_csa_fuse_bwd" + s)(_p(out), _p(gout), _ptr_array(gts), th, tw, len(gts),
                                           B, C, H, W, ct(slope))
    return gts


def refine_frontend_fwd(low, left, right):
    """(low_disp [B,h,w], left [B,C,H,W], right [B,C,H,W]) -> (cat(warped - left, left) [B,2C,H,W],
    disp [B,1,H,W]): refinement.py:80-95 + warp.py:41-64."""
    low = _c(low); left = _c(left, low.dtype); right = _c(right, low.dtype)
    B, h, w = low.shape
    _, C, H, W = left.shape
    concat = np.empty((B, 2 * C, H, W), low.dtype)
    disp = np.empty((B, 1, H, W), low.dtype)
    s, _ = _sfx(low)
    getattr(lib(), "orc_refine_frontend_fwd" + s)(_p(low), _p(left), _p(right), _p(concat), _p(disp),
                                                  B, C, h, w, H, W)
    return concat, disp
