"""aanet_b200 -- B200-native (sm_100a) hot path of AANet stereo matching.

correlation cost volume -> ISA (modulated deformable conv) + CSA fuse -> soft-argmin, behind the
reference's module API.  See DESIGN.md / INTEGRATION.md.
"""
from . import ops  # noqa: F401
from .nets import *  # noqa: F401,F403

__version__ = "0.1.0"
