"""euler_b200 — B200-native (sm_100a) implementation of cgmb/euler's per-timestep fluid solve.

The product is `lib/libeuler_gpu.so` (hand-written CUDA kernels behind the C-ABI declared in
include/euler_gpu.h) and the host C program `bin/euler-gpu` (euler_b200/host/).  This Python
package is only a thin ctypes mirror of that C-ABI for tests and benchmarks; it contains no
compute and no CPU fallback — if the CUDA library is missing, importing `euler_b200.gpu`
raises.
"""
from .scenario import Scenario, resample, synthetic, shipped_text, export_text  # noqa: F401

__all__ = ["Scenario", "resample", "synthetic", "shipped_text", "export_text"]
