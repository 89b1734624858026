"""The C-ABI library loads and exports every symbol include/euler_gpu.h declares (no compute
calls: this file runs without a GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

LIB = os.path.join(ROOT, "euler_b200", "lib", "libeuler_gpu.so")
HEADER = os.path.join(ROOT, "include", "euler_gpu.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(euler_gpu_\w+)\s*\(", src)))


def test_library_is_built():
    assert os.path.exists(LIB), "run `make gpu` (or __graft_entry__.build())"


def test_every_declared_symbol_is_exported():
    lib = ctypes.CDLL(LIB)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libeuler_gpu.so does not export %s" % n


def test_abi_version_and_default_params():
    from euler_b200 import gpu as G
    assert G.abi_version() == 4
    p = G.default_params()
    assert p.pcg_dtype == G.PCG_FP64 and p.pcg_refresh_every == 10
    # the reference's constants (main.c:58-60, 735-736, 838, 849-851)
    assert (p.h, p.rho, p.gravity) == (1.0, 1.0, -10.0)
    assert abs(p.frame_time - 0.1) < 1e-8 and p.max_substeps == 8 and p.cfl_distance == 0.75
    assert p.max_iterations == 100 and p.tol == float(ctypes.c_float(1e-6).value)
    assert p.rng_state == 0x9bd185c449534b91


def test_no_cpu_fallback():
    """Without a CUDA device create() must fail loudly, not fall back to anything."""
    import numpy as np
    from euler_b200 import gpu as G
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    z = np.zeros((8, 8), np.uint8)
    with pytest.raises(G.EulerGpuError) as e:
        G.EulerGpu(8, 8, z, z, z, np.zeros((0, 2), np.float32))
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    """Nothing under euler_b200/ may reference the oracle (it is test infrastructure)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "euler_b200")):
        for f in files:
            if f.endswith((".py", ".c", ".h", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "liboracle" not in src and "euler_oracle.h" not in src, f


def test_pcg_dtype_parameter_validation():
    """pcg_dtype = FP32 is a mode of the fused red-black iteration (single GPU or slabs): every
    other combination is refused before any device work (so this runs without a GPU)."""
    import numpy as np
    from euler_b200 import gpu as G
    z = np.zeros((16, 16), np.uint8)
    m = np.zeros((0, 2), np.float32)
    bad = [dict(pcg_dtype=G.PCG_FP32),                                          # default precon is IC(0)
           dict(pcg_dtype=G.PCG_FP32, precon=G.PRECON_REDBLACK, dot_mode=G.DOT_REFERENCE_ORDER),
           dict(pcg_dtype=G.PCG_FP32, precon=G.PRECON_REDBLACK, stencil_variant=1)]
    for kw in bad:
        with pytest.raises(G.EulerGpuError) as e:
            G.EulerGpu(16, 16, z, z, z, m, **kw)
        assert e.value.code == -4 and "pcg_dtype=FP32" in str(e.value), kw
    for kw, code in ((dict(pcg_dtype=7), -1),
                     (dict(pcg_dtype=G.PCG_FP32, precon=G.PRECON_REDBLACK, pcg_refresh_every=5), -1),
                     (dict(pcg_dtype=G.PCG_FP32, precon=G.PRECON_REDBLACK, pcg_refresh_every=-2), -1)):
        with pytest.raises(G.EulerGpuError) as e:
            G.EulerGpu(16, 16, z, z, z, m, **kw)
        assert e.value.code == code, kw
