"""Swap the B200 hot path into an UNMODIFIED checkout of the reference.

    import sys, types
    sys.path.insert(0, "/path/to/aanet")              # the reference repository
    import aanet_b200.dropin as dropin
    dropin.install()                                  # before `import nets`
    import nets
    model = nets.AANet(192, 0, feature_type='aanet', feature_pyramid_network=True, ...)

`install()` registers this package's modules under the names the reference imports
(`nets.deform_conv`, `nets.deform_conv.deform_conv`, and a stub for the compiled
`nets.deform_conv.deform_conv_cuda` the reference needs at import time, deform_conv.py:9), so that
`nets/deform.py:3`, `nets/feature.py`, `nets/refinement.py` and `thop/profile.py:7` all pick up the
sm_100a operator.  `patch(nets)` then rebinds the hot-path classes that `nets/aanet.py:7-10` imported
by name.  state_dict keys are identical, so reference checkpoints load with strict=True.
"""
import sys
import types

import importlib

from . import nets as _nets

# `nets.deform_conv` the attribute is the *function* (re-exported by the package, as in the
# reference's nets/deform_conv/__init__.py:2-4), so fetch the modules through importlib.
_dc_pkg = importlib.import_module(__package__ + ".nets.deform_conv")
_dc_mod = importlib.import_module(__package__ + ".nets.deform_conv.deform_conv")


def install():
    """Make `from nets.deform_conv import ...` resolve to the B200 operator.  Call before `import nets`."""
    sys.modules.setdefault("nets.deform_conv.deform_conv_cuda", types.ModuleType("deform_conv_cuda"))
    sys.modules["nets.deform_conv"] = _dc_pkg
    sys.modules["nets.deform_conv.deform_conv"] = _dc_mod


def patch(ref_nets):
    """Rebind the hot-path classes inside an imported reference `nets` package."""
    aanet_mod = sys.modules[ref_nets.__name__ + ".aanet"]
    aanet_mod.CostVolume = _nets.CostVolume
    aanet_mod.CostVolumePyramid = _nets.CostVolumePyramid
    aanet_mod.AdaptiveAggregation = _nets.AdaptiveAggregation
    aanet_mod.DisparityEstimation = _nets.DisparityEstimation
    # refinement nets with the fused upsample/warp/error front end (aanet.py:9 imports them by name)
    for attr in ("StereoDRNetRefinement", "HourglassRefinement"):
        if hasattr(aanet_mod, attr):
            setattr(aanet_mod, attr, getattr(_nets, attr))
    deform_mod = sys.modules.get(ref_nets.__name__ + ".deform")
    if deform_mod is not None:      # feature extractor / refinement build DeformConv2d from here
        deform_mod.DeformConv = _nets.DeformConv
        deform_mod.ModulatedDeformConv = _nets.ModulatedDeformConv
        deform_mod.DeformConv2d = _nets.DeformConv2d
        deform_mod.DeformSimpleBottleneck = _nets.DeformSimpleBottleneck
        deform_mod.DeformBottleneck = _nets.DeformBottleneck
    for name in ("resnet", "feature", "refinement"):
        m = sys.modules.get("%s.%s" % (ref_nets.__name__, name))
        if m is not None:
            for attr in ("DeformConv2d", "DeformBottleneck", "DeformSimpleBottleneck"):
                if hasattr(m, attr):
                    setattr(m, attr, getattr(_nets, attr))
    return ref_nets
