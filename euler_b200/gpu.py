"""ctypes mirror of include/euler_gpu.h (the C-ABI of lib/libeuler_gpu.so).

Same names, argument meaning and error behaviour as the C interface; errors become
`EulerGpuError`.  No computation happens here and there is NO fallback: if the CUDA library
cannot be loaded the import fails loudly.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# (EULER_GPU_LIB: an A/B build of the same library, tools only)
LIB_PATH = os.environ.get("EULER_GPU_LIB") or os.path.join(_HERE, "lib", "libeuler_gpu.so")

if not os.path.exists(LIB_PATH):
    raise ImportError("%s is missing — build it with `make gpu` (there is no CPU fallback)" % LIB_PATH)
_L = C.CDLL(LIB_PATH)

PRECON_IC0_WAVEFRONT, PRECON_REDBLACK = 0, 1
MARKERS_REFERENCE, MARKERS_FAST = 0, 1
DOT_TREE, DOT_REFERENCE_ORDER = 0, 1
PCG_FP64, PCG_FP32 = 0, 1
(F_U, F_V, F_UTMP, F_VTMP, F_SOLID, F_SOURCE, F_SINK, F_COUNT, F_PREV_COUNT, F_MARKERS,
 F_PRECON, F_Q, F_ADIAG, F_P, F_R, F_Z, F_S, F_CR, F_CG, F_CB,
 F_R32, F_Z32, F_S32, F_Q32, F_PRECON32) = range(25)
(S_ADVECT_MARKERS, S_REFRESH_COUNTS, S_SOURCES, S_EXTRAPOLATE, S_ADVECT_VELOCITY, S_PROJECT,
 S_BUILD_RHS, S_PRECONDITION, S_APPLY_A, S_PRESSURE_UPDATE, S_EXTRAPOLATE_COLOR,
 S_ADVECT_COLOR, S_FUSED_TAIL) = range(13)

_DTYPES = {F_U: np.float32, F_V: np.float32, F_UTMP: np.float32, F_VTMP: np.float32,
           F_SOLID: np.uint8, F_SOURCE: np.uint8, F_SINK: np.uint8, F_COUNT: np.uint8,
           F_PREV_COUNT: np.uint8, F_PRECON: np.float64, F_Q: np.float64, F_ADIAG: np.int8,
           F_P: np.float64, F_R: np.float64, F_Z: np.float64, F_S: np.float64,
           F_CR: np.float32, F_CG: np.float32, F_CB: np.float32,
           F_R32: np.float32, F_Z32: np.float32, F_S32: np.float32, F_Q32: np.float32,
           F_PRECON32: np.float32}


class Params(C.Structure):
    _fields_ = [("h", C.c_float), ("rho", C.c_float), ("gravity", C.c_float),
                ("frame_time", C.c_float), ("max_substeps", C.c_int), ("cfl_distance", C.c_float),
                ("max_iterations", C.c_int), ("tol", C.c_double),
                ("precon", C.c_int), ("marker_mode", C.c_int), ("dot_mode", C.c_int),
                ("rng_state", C.c_uint64),
                ("device", C.c_int), ("stream", C.c_void_p), ("pcg_check_every", C.c_int), ("stencil_variant", C.c_int),
                ("slab_row0", C.c_int), ("slab_rows", C.c_int), ("rainbow", C.c_int),
                ("pcg_dtype", C.c_int), ("pcg_refresh_every", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("frames", C.c_uint64), ("substeps", C.c_uint64),
                ("solves", C.c_uint64), ("solves_skipped", C.c_uint64),
                ("pcg_iterations", C.c_uint64), ("last_iterations", C.c_int),
                ("last_residual", C.c_double), ("last_dt", C.c_float),
                ("n_markers", C.c_uint64), ("source_exhausted", C.c_int),
                ("rng_state", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("device_bytes", C.c_uint64),
                ("ms_markers", C.c_double), ("ms_grid", C.c_double), ("ms_project", C.c_double),
                ("active_cells", C.c_uint64),
                ("kernel_ms", C.c_double * 24), ("kernel_count", C.c_uint64 * 24),
                ("markers_migrated", C.c_uint64), ("grid_cells", C.c_uint64)]


class Check(C.Structure):
    """euler_check: invariants over the rows the handle owns (additive / max-combinable over slabs)."""
    _fields_ = [("n_markers", C.c_uint64), ("fluid_cells", C.c_uint64), ("count_sum", C.c_uint64),
                ("count_hash", C.c_uint64), ("sum_abs_u", C.c_double), ("sum_abs_v", C.c_double),
                ("sum_p", C.c_double), ("max_abs_div", C.c_double), ("max_abs_u", C.c_double),
                ("max_abs_v", C.c_double)]
    SUMS = ("n_markers", "fluid_cells", "count_sum", "count_hash", "sum_abs_u", "sum_abs_v", "sum_p")
    MAXES = ("max_abs_div", "max_abs_u", "max_abs_v")

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class EulerGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("euler_gpu error %d: %s" % (code, msg))
        self.code = code


_H = C.c_void_p
_L.euler_gpu_last_error.restype = C.c_char_p
_L.euler_gpu_default_params.argtypes = [C.POINTER(Params)]
_L.euler_gpu_create.argtypes = [C.POINTER(_H), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_size_t, C.POINTER(Params)]
_L.euler_gpu_destroy.argtypes = [_H]
_L.euler_gpu_reinit.argtypes = [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64]
_L.euler_gpu_step_frame.argtypes = [_H, C.POINTER(C.c_int)]
_L.euler_gpu_calculate_timestep.argtypes = [_H, C.c_float, C.POINTER(C.c_float)]
_L.euler_gpu_substep.argtypes = [_H, C.c_float]
_L.euler_gpu_run_stage.argtypes = [_H, C.c_int, C.c_float]
_L.euler_gpu_read_marker_count.argtypes = [_H, C.c_void_p]
_L.euler_gpu_get.argtypes = [_H, C.c_int, C.c_void_p, C.c_size_t]
_L.euler_gpu_read_window.argtypes = [_H, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
_L.euler_gpu_set.argtypes = [_H, C.c_int, C.c_void_p, C.c_size_t]
_L.euler_gpu_set_rng_state.argtypes = [_H, C.c_uint64]
_L.euler_gpu_colorize.argtypes = [_H]
_L.euler_gpu_set_source_exhausted.argtypes = [_H, C.c_int]
_L.euler_gpu_set_frame_count.argtypes = [_H, C.c_uint64]
_L.euler_gpu_set_max_iterations.argtypes = [_H, C.c_int]
_L.euler_gpu_stats.argtypes = [_H, C.POINTER(Stats)]
_L.euler_gpu_check.argtypes = [_H, C.POINTER(Check)]
_L.euler_gpu_set_profiling.argtypes = [_H, C.c_int]
_L.euler_gpu_trace_read.argtypes = [_H, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
_L.euler_gpu_synchronize.argtypes = [_H]
_L.euler_gpu_reset_profile.argtypes = [_H]
_L.euler_gpu_kernel_class_name.restype = C.c_char_p
_L.euler_gpu_kernel_class_name.argtypes = [C.c_int]
_L.euler_gpu_stream.restype = C.c_void_p
_L.euler_gpu_stream.argtypes = [_H]
_L.euler_gpu_pcg_iterations.argtypes = [_H, C.c_int]
_L.euler_gpu_comm_unique_id.argtypes = [C.c_void_p]
_L.euler_gpu_comm_init.argtypes = [_H, C.c_int, C.c_int, C.c_void_p]
_L.euler_gpu_comm_p2p_export.argtypes = [_H, C.c_void_p]
_L.euler_gpu_comm_p2p_import.argtypes = [_H, C.c_void_p]
_L.euler_gpu_slab_partition.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
_L.euler_gpu_slab_partition_weighted.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]


def _ck(rc):
    if rc != 0:
        raise EulerGpuError(rc, _L.euler_gpu_last_error().decode(errors="replace"))


def default_params():
    p = Params()
    _ck(_L.euler_gpu_default_params(C.byref(p)))
    return p


def slab_partition(global_ny, n_ranks, rank):
    """(row0, rows) of `rank` in the balanced row-slab split (euler_gpu_slab_partition)."""
    a, b = C.c_int(0), C.c_int(0)
    _ck(_L.euler_gpu_slab_partition(global_ny, n_ranks, rank, C.byref(a), C.byref(b)))
    return a.value, b.value


def slab_partition_weighted(row_weight, n_ranks, rank):
    w = np.ascontiguousarray(row_weight, dtype=np.uint64)
    a, b = C.c_int(0), C.c_int(0)
    _ck(_L.euler_gpu_slab_partition_weighted(w.ctypes.data, len(w), n_ranks, rank, C.byref(a), C.byref(b)))
    return a.value, b.value


def comm_unique_id():
    buf = C.create_string_buffer(128)
    _ck(_L.euler_gpu_comm_unique_id(buf))
    return buf.raw


def abi_version():
    return int(_L.euler_gpu_abi_version())


class EulerGpu:
    """One simulation handle (see include/euler_gpu.h for the meaning of every call)."""

    def __init__(self, nx, ny, solid, source, sink, markers, params=None, **overrides):
        p = params or default_params()
        for k, v in overrides.items():
            if not hasattr(p, k):
                raise TypeError("unknown parameter %r" % k)
            setattr(p, k, v)
        self.params = p
        self.nx, self.ny = nx, ny
        solid = np.ascontiguousarray(solid, dtype=np.uint8)
        source = np.ascontiguousarray(source, dtype=np.uint8)
        sink = np.ascontiguousarray(sink, dtype=np.uint8)
        markers = np.ascontiguousarray(markers, dtype=np.float32).reshape(-1, 2)
        self._h = _H()
        _ck(_L.euler_gpu_create(C.byref(self._h), nx, ny, solid.ctypes.data, source.ctypes.data,
                                sink.ctypes.data, markers.ctypes.data, len(markers), C.byref(p)))

    @classmethod
    def from_scenario(cls, scn, **overrides):
        overrides.setdefault("rng_state", scn.rng_state)
        return cls(scn.nx, scn.ny, scn.solid, scn.source, scn.sink, scn.markers, **overrides)

    def reinit(self, solid, source, sink, markers, rng_state):
        """sim_init() again on this handle (euler_gpu_reinit): host arrays in, no allocation."""
        solid = np.ascontiguousarray(solid, dtype=np.uint8)
        source = np.ascontiguousarray(source, dtype=np.uint8)
        sink = np.ascontiguousarray(sink, dtype=np.uint8)
        markers = np.ascontiguousarray(markers, dtype=np.float32).reshape(-1, 2)
        for a in (solid, source, sink):
            if a.shape != (self.ny, self.nx):
                raise ValueError("shape %r != %r" % (a.shape, (self.ny, self.nx)))
        _ck(_L.euler_gpu_reinit(self._h, solid.ctypes.data, source.ctypes.data, sink.ctypes.data,
                                markers.ctypes.data, len(markers), int(rng_state)))

    def close(self):
        if getattr(self, "_h", None):
            _L.euler_gpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self): return self
    def __exit__(self, *a): self.close()

    def comm_init(self, rank, n_ranks, unique_id):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        _ck(_L.euler_gpu_comm_init(self._h, rank, n_ranks, buf))

    def comm_p2p_export(self):
        buf = C.create_string_buffer(256)
        _ck(_L.euler_gpu_comm_p2p_export(self._h, buf))
        return buf.raw

    def comm_p2p_import(self, blobs):
        data = b"".join(blobs)
        buf = C.create_string_buffer(data, len(data))
        _ck(_L.euler_gpu_comm_p2p_import(self._h, buf))

    def step_frame(self):
        n = C.c_int(0)
        _ck(_L.euler_gpu_step_frame(self._h, C.byref(n)))
        return n.value

    def calculate_timestep(self, frame_time=0.1):
        dt = C.c_float(0)
        _ck(_L.euler_gpu_calculate_timestep(self._h, np.float32(frame_time), C.byref(dt)))
        return float(dt.value)

    def substep(self, dt): _ck(_L.euler_gpu_substep(self._h, np.float32(dt)))
    def run_stage(self, stage, dt=0.0): _ck(_L.euler_gpu_run_stage(self._h, stage, np.float32(dt)))
    def pcg_iterations(self, n): _ck(_L.euler_gpu_pcg_iterations(self._h, n))
    def synchronize(self): _ck(_L.euler_gpu_synchronize(self._h))
    def trace_read(self, max_slots=65536):
        """Timeline slots recorded since the last call (EULER_TRACE=<slots>): array [n, 16] of uint64."""
        out = np.zeros((max_slots, 16), dtype=np.uint64)
        n = C.c_size_t(0)
        _ck(_L.euler_gpu_trace_read(self._h, out.ctypes.data_as(C.c_void_p), C.c_size_t(max_slots), C.byref(n)))
        self.trace_blocks = out[n.value:n.value + 1024].reshape(-1, 4) if max_slots >= n.value + 1024 else None
        return out[:n.value]
    def set_profiling(self, on): _ck(_L.euler_gpu_set_profiling(self._h, 1 if on else 0))
    def reset_profile(self): _ck(_L.euler_gpu_reset_profile(self._h))

    def kernel_profile(self):
        """{class name: (total ms, timed launch groups)} accumulated while profiling was on."""
        st = self.stats()
        out = {}
        for i in range(24):
            name = _L.euler_gpu_kernel_class_name(i)
            if name and st.kernel_count[i]:
                out[name.decode()] = (float(st.kernel_ms[i]), int(st.kernel_count[i]))
        return out

    def colorize(self): _ck(_L.euler_gpu_colorize(self._h))
    def set_rng_state(self, s): _ck(_L.euler_gpu_set_rng_state(self._h, s))
    def set_frame_count(self, n): _ck(_L.euler_gpu_set_frame_count(self._h, int(n)))
    def set_max_iterations(self, n): _ck(_L.euler_gpu_set_max_iterations(self._h, int(n)))
    def set_source_exhausted(self, e): _ck(_L.euler_gpu_set_source_exhausted(self._h, 1 if e else 0))

    @property
    def stream(self): return _L.euler_gpu_stream(self._h)

    def stats(self):
        s = Stats()
        _ck(_L.euler_gpu_stats(self._h, C.byref(s)))
        return s

    def check(self):
        c = Check()
        _ck(_L.euler_gpu_check(self._h, C.byref(c)))
        return c

    def read_marker_count(self, out=None):
        if out is None:
            out = np.empty((self.ny, self.nx), dtype=np.uint8)
        _ck(_L.euler_gpu_read_marker_count(self._h, out.ctypes.data))
        return out

    def read_window(self, field, x0, y0, w, h, out):
        """Rectangle [x0,x0+w) x [y0,y0+h) of a plane into the same place of a global-shaped array."""
        if out.shape != (self.ny, self.nx) or out.dtype != _DTYPES[field] or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous (ny, nx) array of the field's dtype")
        _ck(_L.euler_gpu_read_window(self._h, field, x0, y0, w, h, out.ctypes.data))
        return out

    def get(self, field):
        if field == F_MARKERS:
            n = int(self.stats().n_markers)
            out = np.empty((n, 2), dtype=np.float32)
            _ck(_L.euler_gpu_get(self._h, field, out.ctypes.data, out.nbytes))
            return out
        out = np.empty((self.ny, self.nx), dtype=_DTYPES[field])
        _ck(_L.euler_gpu_get(self._h, field, out.ctypes.data, out.nbytes))
        return out

    def set(self, field, arr):
        if field == F_MARKERS:
            a = np.ascontiguousarray(arr, dtype=np.float32).reshape(-1, 2)
        else:
            a = np.ascontiguousarray(arr, dtype=_DTYPES[field])
            if a.shape != (self.ny, self.nx):
                raise ValueError("shape %r != %r" % (a.shape, (self.ny, self.nx)))
        _ck(_L.euler_gpu_set(self._h, field, a.ctypes.data, a.nbytes))
