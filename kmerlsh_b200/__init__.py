"""kmerlsh_b200 — B200-native (sm_100a) implementation of kmerLSH's mode-C clustering hot path.

The product is `libklsh.so` (hand-written CUDA kernels behind the C ABI in include/klsh.h) and the
mode-C command line `kmerLSH_b200`.  This package is the thin ctypes mirror of the reference's
`Cluster(...)` seam used by the tests and the benchmark; it never computes on the CPU and raises
if the CUDA library is missing or no B200 is present.
"""
from .api import (  # noqa: F401
    Context,
    KlshError,
    IterStats,
    Cluster,
    lib_path,
    load_library,
    nccl_unique_id,
)
