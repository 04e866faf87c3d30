"""daliti_b200 -- B200-native (sm_100a CUDA) scan-to-map IEKF measurement update of DaLiTI's eskf_lio.

The product is the in-tree shared library ``daliti_b200/lib/libdaliti_b200.so`` (hand-written
CUDA kernels behind the C ABI of ``include/daliti_b200.h`` / ``include/daliti_b200_lio.h``).
This package is a thin ctypes loader for tests and benchmarks; it has no CPU fallback and
raises if the library is missing or no CUDA device is usable.
"""
from .binding import (  # noqa: F401
    DltConfig,
    DltError,
    Measurement,
    ScanToMap,
    default_library_path,
    load_library,
)
from .lio import LaserMapping, LioConfig, LioScanOut, LioThermal  # noqa: F401
