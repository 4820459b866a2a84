"""mplambda_b200 -- B200-native validity checking for mplambda-style planners.

The product is libmplb.so (CUDA, sm_100a) behind the C ABI in include/mplb.h; this package is
the thin host-side mirror of the reference's scenario interface over that ABI.
"""
from .scenario import FetchScenario, SE3RigidBodyScenario, load_collada  # noqa: F401
from .nn import Nigh  # noqa: F401
from ._capi import MplbError, lib, lib_path  # noqa: F401
