"""Drop-in for the reference's nets/deform_conv package (same public names, __init__.py:2-9)."""
from .deform_conv import (DeformConv, DeformConvPack, ModulatedDeformConv, ModulatedDeformConvPack,
                          deform_conv, modulated_deform_conv)

__all__ = ['DeformConv', 'DeformConvPack', 'ModulatedDeformConv', 'ModulatedDeformConvPack',
           'deform_conv', 'modulated_deform_conv']
