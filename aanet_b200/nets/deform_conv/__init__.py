"""Package twin of the reference's nets/deform_conv: exposes the operator classes and the two functional
entry points under the names `nets/deform.py`, `nets/feature.py` and `thop/profile.py` import.

As in the reference, the package attribute `deform_conv` ends up being the FUNCTION (it shadows the
sub-module of the same name); code that needs the module goes through importlib (see dropin.py).
"""
import importlib as _importlib

_impl = _importlib.import_module(__name__ + ".deform_conv")

__all__ = sorted(name for name in vars(_impl)
                 if name in {"deform_conv", "modulated_deform_conv"} or
                 (name.startswith(("DeformConv", "ModulatedDeformConv")) and not name.endswith("Function")))
globals().update({name: getattr(_impl, name) for name in __all__})
