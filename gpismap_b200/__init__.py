"""gpismap_b200 — B200-native hot path of GPisMap (batched leaf-GP train + SDF/gradient/variance query).

The product is the C ABI in include/gpis_b200.h (gpismap_b200/libgpis_b200.so, hand-written sm_100a
kernels under gpismap_b200/csrc/) and the drop-in C++ classes GPisMap / GPisMap3 in include/gpismap/
(gpismap_b200/host/). The Python modules here are bindings used by tests and bench.py.
"""
from . import cabi  # noqa: F401
