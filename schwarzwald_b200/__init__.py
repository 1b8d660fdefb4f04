"""schwarzwald_b200 — B200 (sm_100a) implementation of Schwarzwald's tiler compute core.

The product is the CUDA library ``libswgpu.so`` (C ABI in ``include/swgpu.h``).  This package holds
its sources (``csrc/``), the C++ adapter that mirrors the reference's ``TilingAlgorithmBase``
(``host/``) and a thin ctypes mirror used by the tests, the benchmark and the multi-GPU driver.
There is no CPU fallback: importing works anywhere, every compute call needs the library and a GPU.
"""
from .tiler import (  # noqa: F401
    ACCURATE,
    FAST,
    GRID_CENTER,
    JITTERED,
    MIN_DISTANCE,
    RANDOM_GRID,
    GpuTiler,
    MultiGpuTiler,
    SwgpuError,
    TileResult,
    cubic_bounds,
    cubic_bounds_at_origin,
    node_name,
    spacing_from_diagonal_fraction,
)
from .native import library_path, load_library  # noqa: F401
