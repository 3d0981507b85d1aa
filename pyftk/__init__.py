"""pyftk -- the reference's Python module name (python/pyftk.cpp), served by the B200 engine.

`import pyftk` gives the same three sub-modules the reference's pybind11 module defines:

    pyftk.trackers.track_critical_points_2d_scalar(array)      python/pyftk.cpp:92-142
    pyftk.extractors.extract_critical_points_2d_scalar(array)  python/pyftk.cpp:15-52
    pyftk.extractors.extract_critical_points_2d_vector(array)  python/pyftk.cpp:54-90
    pyftk.synthesizers.{spiral_woven, double_gyre_flow, moving_extremum}   python/pyftk.cpp:144-179

with the reference's array contracts (numpy memory reinterpreted dim-0-fastest, not transposed).  Everything runs in
libftkb200.so on the GPU (ftk_b200/); there is no CPU path, so importing works anywhere but calling needs a B200.
"""
from ftk_b200 import extractors, synthesizers, trackers  # noqa: F401
from ftk_b200 import __version__  # noqa: F401
