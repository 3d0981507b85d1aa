"""ftk_b200 -- B200-native critical-point extraction and tracking behind FTK's operator surface.

The product is the C-ABI library libftkb200.so (include/ftkb200.h; CUDA kernels for sm_100a in
ftk_b200/csrc).  This package is the Python-facing mirror of the reference's pyftk module
(python/pyftk.cpp): `trackers`, `extractors`, `synthesizers`, plus the tracker classes themselves.
Nothing here computes on the CPU: without the built library and a B200 the calls fail loudly.
"""
from . import _lib
from .tracker import (Lattice, critical_point_tracker_2d_regular, critical_point_tracker_3d_regular,  # noqa: F401
                      critical_point_type_to_string, make_tracker, track, SOURCE_NONE, SOURCE_GIVEN, SOURCE_DERIVED)
from . import trackers, extractors, synthesizers  # noqa: F401
from .curves import CurveSet  # noqa: F401

__version__ = "0.1.0"
lattice = Lattice
