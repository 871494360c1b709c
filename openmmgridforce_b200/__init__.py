"""openmmgridforce_b200 — B200 (sm_100a) GridForce evaluation path.

The product is ``lib/libgridforce_b200.so`` (hand-written CUDA behind the C ABI declared in
``include/gridforce_b200.h``) plus the OpenMM platform plugin in ``plugin/``. This Python package is the
ctypes binding the tests and ``bench.py`` drive that C ABI with; it contains no arithmetic of its own and no
fallback: if the shared library is missing, or there is no sm_100 GPU, calls raise.
"""
from .capi import (  # noqa: F401
    GridForceB200Error,
    Device,
    Grid,
    Kernel,
    FORCE_F64_STORE,
    FORCE_F64_ADD,
    FORCE_FIXED_ADD,
    FORCE_F32_STORE,
    Graph,
    Comm,
    Multi,
    host_register,
    DeviceArrayView,
    host_unregister,
    PRECISION_MIXED,
    PRECISION_DOUBLE,
    MAX_GRIDS,
    LAYOUT_AUTO,
    LAYOUT_CELLS,
    LAYOUT_ROWS,
    LAYOUT_PAIRS,
    LAYOUT_BSPLINE,
    LAYOUT_POINTS,
    LAYOUT_HERMITE,
    LAYOUT_BSPLINE_POINTS,
    LAYOUT_NAMES,
    library_path,
    load_library,
    launch_count,
    read_grid_file,
    write_grid_file,
)

__version__ = "0.2.0"
