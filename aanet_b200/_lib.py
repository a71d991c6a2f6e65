"""ctypes binding of libaanet_b200.so (the C-ABI declared in include/aanet_b200.h).

There is no fallback: if the shared library is missing the import of any op raises, and every op
raises NotImplementedError for non-CUDA tensors exactly like the reference op
(nets/deform_conv/deform_conv.py:135-136).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# AANET_B200_LIB: load another build of the same library (A/B timing of kernel variants in one process tree)
LIB_PATH = os.environ.get("AANET_B200_LIB") or os.path.join(_HERE, "lib", "libaanet_b200.so")
ABI_VERSION = 2

_vp, _i, _f, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t

# symbol -> (restype, argtypes); must list every function include/aanet_b200.h declares
SIGNATURES = {
    "aanet_abi_version": (_i, []),
    "aanet_status_string": (ctypes.c_char_p, [_i]),
    "aanet_last_cuda_error": (ctypes.c_char_p, []),
    "aanet_corr_fwd": (_i, [_vp, _vp, _vp] + [_i] * 5 + [_vp]),
    "aanet_corr_fwd_bf16": (_i, [_vp, _vp, _vp] + [_i] * 5 + [_vp]),
    "aanet_corr_fwd_nhwc": (_i, [_vp, _vp, _vp] + [_i] * 5 + [_vp]),
    "aanet_corr_bwd": (_i, [_vp] * 5 + [_i] * 5 + [_vp]),
    "aanet_cost5d_fwd": (_i, [_vp, _vp, _vp] + [_i] * 6 + [_vp]),
    "aanet_cost5d_bwd": (_i, [_vp, _vp, _vp] + [_i] * 6 + [_vp]),
    "aanet_softargmin_fwd": (_i, [_vp, _vp] + [_i] * 5 + [_vp]),
    "aanet_softargmin_bwd": (_i, [_vp, _vp, _vp] + [_i] * 5 + [_vp]),
    "aanet_refine_frontend_fwd": (_i, [_vp] * 5 + [_i] * 6 + [_vp]),
    "aanet_mdcn_workspace_bytes": (_sz, [_i] * 13),
    "aanet_mdcn_fwd": (_i, [_vp] * 6 + [_i] * 12 + [_vp, _vp, _i, _vp, _sz, _vp]),
    "aanet_mdcn_bwd": (_i, [_vp] * 10 + [_i] * 12 + [_vp, _sz, _vp]),
    "aanet_conv2d_workspace_bytes": (_sz, [_i] * 11),
    "aanet_conv2d_fwd": (_i, [_vp] * 6 + [_i, _f, _vp] + [_i] * 11 + [_vp, _sz, _vp]),
    "aanet_conv_wpack_bytes": (_sz, [_i] * 6),
    "aanet_conv_pack_weights": (_i, [_vp, _vp] + [_i] * 6 + [_vp]),
    "aanet_nchw_to_nhwc": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "aanet_nhwc_to_nchw": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "aanet_conv_batch_nhwc": (_i, [_vp, _i, _i, _i, _vp]),
    "aanet_conv_tail_supported": (_i, [_vp, _i]),
    "aanet_csa_fuse_nhwc": (_i, [_vp, _vp, _vp, _i, _vp] + [_i] * 4 + [_f, _vp]),
    "aanet_csa_conv1_nhwc": (_i, [_vp, _vp, _vp, _i, _f, _vp, _vp, _vp, _vp, _vp, _i, _vp] + [_i] * 5 + [_vp]),
    "aanet_csa_fuse_fwd": (_i, [_vp, _vp, _vp, _i, _vp] + [_i] * 4 + [_f, _vp]),
    "aanet_csa_fuse_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i] + [_i] * 4 + [_f, _vp]),
}



class ConvDesc(ctypes.Structure):
    """Mirror of `aanet_conv_desc` (include/aanet_b200.h)."""
    _fields_ = [("x", _vp), ("wpack", _vp), ("bias", _vp), ("scale", _vp), ("shift", _vp), ("residual", _vp),
                ("out", _vp), ("offmask", _vp), ("om_channels", _i),
                ("B", _i), ("Cin", _i), ("H", _i), ("W", _i), ("Cout", _i), ("kh", _i), ("kw", _i),
                ("stride", _i), ("pad", _i), ("dil", _i), ("groups", _i), ("dg", _i),
                ("act", _i), ("slope", _f), ("n_offset_ch", _i), ("mask_scale", _f), ("out_nchw", _i),
                ("om_nchw", _i),
                ("tail_wpack", _vp), ("tail_scale", _vp), ("tail_shift", _vp), ("tail_residual", _vp),
                ("tail_cout", _i), ("tail_act", _i)]


_lib = None


class AanetError(RuntimeError):
    """Non-zero aanet_status from the C-ABI (the reference raises RuntimeError via TORCH_CHECK)."""


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "aanet_b200: %s is missing -- build it with `python -m aanet_b200.build` "
            "(or __graft_entry__.build()). There is no CPU or PyTorch fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)     # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.aanet_abi_version() != ABI_VERSION:
        raise ImportError("aanet_b200: ABI version mismatch (library %d, binding %d)"
                          % (lib.aanet_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        lib = load()
        msg = lib.aanet_status_string(status).decode()
        if status == 5:
            msg += ": " + lib.aanet_last_cuda_error().decode()
        raise AanetError("%s failed: %s (status %d)" % (what, msg, status))
