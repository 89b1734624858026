"""oracle/oracle.py — TEST INFRASTRUCTURE ONLY.

ctypes bindings for (a) `liboracle.so`, this repo's CPU restatement of the reference's
fluid solve, and (b) `oracle/_ref/libeuler_ref_<X>x<Y>[_fast].so`, the UNMODIFIED reference
(cgmb/euler main.c) compiled in place by oracle/build_ref.sh.  Only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module; nothing
under euler_b200/ does.
"""
import ctypes as C
import os
import threading
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

P, U, V = 0, 1, 2
PRECON_IC0, PRECON_REDBLACK = 0, 1
PCG_FP64, PCG_FP32 = 0, 1


class _Vec2(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float)]


class _Sim(C.Structure):
    _fields_ = [
        ("nx", C.c_int), ("ny", C.c_int),
        ("h", C.c_float), ("rho", C.c_float), ("gravity", C.c_float),
        ("max_iterations", C.c_int), ("tol", C.c_double),
        ("precon_mode", C.c_int), ("quirk_marker_dt_leak", C.c_int),
        ("u", C.POINTER(C.c_float)), ("v", C.POINTER(C.c_float)),
        ("utmp", C.POINTER(C.c_float)), ("vtmp", C.POINTER(C.c_float)),
        ("solid", C.POINTER(C.c_uint8)), ("source", C.POINTER(C.c_uint8)),
        ("sink", C.POINTER(C.c_uint8)), ("count", C.POINTER(C.c_uint8)),
        ("prev_count", C.POINTER(C.c_uint8)),
        ("markers", C.POINTER(_Vec2)),
        ("n_markers", C.c_size_t), ("max_markers", C.c_size_t),
        ("source_exhausted", C.c_int),
        ("rng_state", C.c_uint64), ("rng_draws", C.c_uint64),
        ("adiag", C.POINTER(C.c_int8)),
        ("precon", C.POINTER(C.c_double)), ("q", C.POINTER(C.c_double)),
        ("b", C.POINTER(C.c_double)), ("p", C.POINTER(C.c_double)),
        ("r", C.POINTER(C.c_double)), ("z", C.POINTER(C.c_double)),
        ("s", C.POINTER(C.c_double)),
        ("rainbow", C.c_int),
        ("cr", C.POINTER(C.c_float)), ("cg", C.POINTER(C.c_float)), ("cb", C.POINTER(C.c_float)),
        ("crtmp", C.POINTER(C.c_float)), ("cgtmp", C.POINTER(C.c_float)), ("cbtmp", C.POINTER(C.c_float)),
        ("frame_count", C.c_uint16),
        ("last_iterations", C.c_int), ("last_solve_skipped", C.c_int),
        ("last_residual", C.c_double),
        ("total_iterations", C.c_long), ("total_substeps", C.c_long),
        ("total_solves", C.c_long), ("last_dt", C.c_float),
        ("pcg_dtype", C.c_int), ("refresh_every", C.c_int),
        ("r32", C.POINTER(C.c_float)), ("z32", C.POINTER(C.c_float)),
        ("s32", C.POINTER(C.c_float)), ("q32", C.POINTER(C.c_float)),
        ("as32", C.POINTER(C.c_float)), ("pc32", C.POINTER(C.c_float)),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/liboracle.so missing: run `make -C oracle oracle`")
        L = C.CDLL(path)
        SP = C.POINTER(_Sim)
        FP, DP = C.POINTER(C.c_float), C.POINTER(C.c_double)
        L.orc_create.restype = SP; L.orc_create.argtypes = [C.c_int, C.c_int]
        L.orc_destroy.argtypes = [SP]
        L.orc_init_from_text.argtypes = [SP, C.c_char_p, C.c_int]
        L.orc_calculate_timestep.restype = C.c_float
        L.orc_calculate_timestep.argtypes = [SP, C.c_float]
        L.orc_advect_markers.argtypes = [SP, C.c_float]
        L.orc_refresh_marker_counts.argtypes = [SP]
        L.orc_update_fluid_sources.argtypes = [SP]
        L.orc_extrapolate.argtypes = [SP, FP, C.c_int]
        L.orc_zero_bounds.argtypes = [SP, FP, C.c_int]
        L.orc_advect_u.argtypes = [SP, FP, FP, C.c_float, FP]
        L.orc_advect_v.argtypes = [SP, FP, FP, C.c_float, FP]
        L.orc_apply_body_forces.argtypes = [SP, FP, C.c_float]
        L.orc_project.argtypes = [SP, C.c_float, FP, FP, FP, FP]
        L.orc_substep.argtypes = [SP, C.c_float]
        L.orc_step_frame.restype = C.c_int; L.orc_step_frame.argtypes = [SP]
        L.orc_build_rhs.argtypes = [SP, C.c_float, FP, FP]
        L.orc_apply_preconditioner.argtypes = [SP, DP, DP]
        L.orc_apply_a.argtypes = [SP, DP, DP]
        L.orc_dot.restype = C.c_double; L.orc_dot.argtypes = [SP, DP, DP]
        L.orc_inf_norm.restype = C.c_double; L.orc_inf_norm.argtypes = [SP, DP]
        L.orc_all_zero.restype = C.c_int; L.orc_all_zero.argtypes = [SP, DP]
        L.orc_pressure_update.argtypes = [SP, C.c_float, FP, FP, FP, FP]
        L.orc_interpolate.restype = C.c_float
        L.orc_interpolate.argtypes = [SP, FP, C.c_float, C.c_float, C.c_int]
        L.orc_randf.restype = C.c_float; L.orc_randf.argtypes = [SP]
        L.orc_hsv_basis.restype = C.c_float; L.orc_hsv_basis.argtypes = [C.c_float]
        L.orc_colorize.argtypes = [SP]
        L.orc_advect_p.argtypes = [SP, FP, FP, FP, C.c_float, FP]
        L.orc_rb_build32.argtypes = [SP]
        L.orc_rb_apply32.argtypes = [SP, FP, FP]
        L.orc_apply_a32.argtypes = [SP, FP, FP]
        L.orc_fnv1a.restype = C.c_uint64
        L.orc_fnv1a.argtypes = [C.POINTER(C.c_uint8), C.c_size_t]
        _lib = L
    return _lib


_F32 = ("u", "v", "utmp", "vtmp", "cr", "cg", "cb", "crtmp", "cgtmp", "cbtmp",
        "r32", "z32", "s32", "q32", "as32", "pc32")
_U8 = ("solid", "source", "sink", "count", "prev_count")
_F64 = ("precon", "q", "b", "p", "r", "z", "s")


class Oracle:
    """The restatement.  Planes are exposed as numpy views [ny, nx] on the C arrays."""

    def __init__(self, nx, ny, text=None, rainbow=False):
        self.L = lib()
        self.ptr = self.L.orc_create(nx, ny)
        self.c = self.ptr.contents
        self.c.rainbow = 1 if rainbow else 0
        self.nx, self.ny = nx, ny
        shape = (ny, nx)
        for name in _F32 + _U8 + _F64 + ("adiag",):
            setattr(self, name, np.ctypeslib.as_array(getattr(self.c, name), shape=shape))
        self._markers = np.ctypeslib.as_array(
            C.cast(self.c.markers, C.POINTER(C.c_float)), shape=(4 * nx * ny, 2))
        if text is not None:
            self.init_from_text(text)

    def __del__(self):
        try:
            self.L.orc_destroy(self.ptr)
        except Exception:
            pass

    # ---- state
    @property
    def n_markers(self):
        return int(self.c.n_markers)

    @property
    def markers(self):
        return self._markers[: self.n_markers]

    def set_markers(self, m):
        m = np.ascontiguousarray(m, dtype=np.float32).reshape(-1, 2)
        self._markers[: len(m)] = m
        self.c.n_markers = len(m)

    def fptr(self, a):
        return a.ctypes.data_as(C.POINTER(C.c_float))

    def dptr(self, a):
        return a.ctypes.data_as(C.POINTER(C.c_double))

    def init_from_text(self, text):
        if isinstance(text, str):
            text = text.encode()
        self.L.orc_init_from_text(self.ptr, text, len(text))

    # ---- stages
    def calculate_timestep(self, frame_time=0.1):
        return float(self.L.orc_calculate_timestep(self.ptr, np.float32(frame_time)))

    def advect_markers(self, dt): self.L.orc_advect_markers(self.ptr, np.float32(dt))
    def refresh_marker_counts(self): self.L.orc_refresh_marker_counts(self.ptr)
    def update_fluid_sources(self): self.L.orc_update_fluid_sources(self.ptr)
    def extrapolate(self, q, t): self.L.orc_extrapolate(self.ptr, self.fptr(q), t)
    def zero_bounds(self, q, t): self.L.orc_zero_bounds(self.ptr, self.fptr(q), t)

    def advect_u(self, dt):
        self.L.orc_advect_u(self.ptr, self.fptr(self.u), self.fptr(self.v), np.float32(dt), self.fptr(self.utmp))

    def advect_v(self, dt):
        self.L.orc_advect_v(self.ptr, self.fptr(self.u), self.fptr(self.v), np.float32(dt), self.fptr(self.vtmp))

    def apply_body_forces(self, dt): self.L.orc_apply_body_forces(self.ptr, self.fptr(self.vtmp), np.float32(dt))

    def project(self, dt):
        self.L.orc_project(self.ptr, np.float32(dt), self.fptr(self.utmp), self.fptr(self.vtmp),
                           self.fptr(self.u), self.fptr(self.v))

    def build_rhs(self, dt):
        self.L.orc_build_rhs(self.ptr, np.float32(dt), self.fptr(self.utmp), self.fptr(self.vtmp))

    def apply_preconditioner(self, r, z): self.L.orc_apply_preconditioner(self.ptr, self.dptr(r), self.dptr(z))
    def apply_a(self, s, out): self.L.orc_apply_a(self.ptr, self.dptr(s), self.dptr(out))
    def dot(self, a, b): return float(self.L.orc_dot(self.ptr, self.dptr(a), self.dptr(b)))
    def inf_norm(self, r): return float(self.L.orc_inf_norm(self.ptr, self.dptr(r)))

    # mixed-precision mirror (fp32 planes r32, z32, s32, q32, as32, pc32; see euler_oracle.h)
    def rb_build32(self): self.L.orc_rb_build32(self.ptr)
    def rb_apply32(self, r, z): self.L.orc_rb_apply32(self.ptr, self.fptr(r), self.fptr(z))
    def apply_a32(self, s, out): self.L.orc_apply_a32(self.ptr, self.fptr(s), self.fptr(out))

    def pressure_update(self, dt):
        self.L.orc_pressure_update(self.ptr, np.float32(dt), self.fptr(self.utmp), self.fptr(self.vtmp),
                                   self.fptr(self.u), self.fptr(self.v))

    def colorize(self): self.L.orc_colorize(self.ptr)

    def advect_p(self, q, dt, out):
        self.L.orc_advect_p(self.ptr, self.fptr(q), self.fptr(self.u), self.fptr(self.v), np.float32(dt), self.fptr(out))

    def substep(self, dt): self.L.orc_substep(self.ptr, np.float32(dt))
    def step_frame(self): return int(self.L.orc_step_frame(self.ptr))

    def interpolate(self, q, ix, iy, t):
        return float(self.L.orc_interpolate(self.ptr, self.fptr(q), np.float32(ix), np.float32(iy), t))

    def fnv_count(self):
        return int(self.L.orc_fnv1a(self.count.ctypes.data_as(C.POINTER(C.c_uint8)), self.count.size))


def fnv1a(arr):
    a = np.ascontiguousarray(arr, dtype=np.uint8)
    return int(lib().orc_fnv1a(a.ctypes.data_as(C.POINTER(C.c_uint8)), a.size))


# --------------------------------------------------------------------------- reference

def ref_available(nx, ny, fast=False):
    return os.path.exists(_ref_path(nx, ny, fast))


def _ref_path(nx, ny, fast=False):
    return os.path.join(REF_DIR, "libeuler_ref_%dx%d%s.so" % (nx, ny, "_fast" if fast else ""))


class _Args(C.Structure):                    # args_t, reference main.c:52-55
    _fields_ = [("scenario_file", C.c_char_p), ("rainbow", C.c_bool)]


def call_with_big_stack(fn, *args, stack_mb=None):
    """project() keeps five double[Y][X] VLAs on the stack (main.c:716,739-745): run the call on
    a thread whose stack is large enough."""
    out = {}

    def run():
        try:
            out["v"] = fn(*args)
        except BaseException as e:     # pragma: no cover
            out["e"] = e
    old = threading.stack_size()
    threading.stack_size((stack_mb or 64) * 1024 * 1024)
    try:
        t = threading.Thread(target=run)
        t.start()
        t.join()
    finally:
        threading.stack_size(old)
    if "e" in out:
        raise out["e"]
    return out.get("v")


class Reference:
    """The unmodified reference, one .so per compile-time grid size.  Each instance loads a
    PRIVATE copy of the library (its state is file-scope globals), so several can coexist."""
    _n = 0

    def __init__(self, nx, ny, fast=False):
        import shutil, tempfile
        path = _ref_path(nx, ny, fast)
        if not os.path.exists(path):
            raise RuntimeError("%s missing: run oracle/build_ref.sh %d %d" % (path, nx, ny))
        # dlopen caches by path: copy to a unique temp name to get private globals
        Reference._n += 1
        self._tmp = tempfile.NamedTemporaryFile(suffix="_%d.so" % Reference._n, delete=False)
        self._tmp.close()
        shutil.copyfile(path, self._tmp.name)
        self.L = C.CDLL(self._tmp.name)
        os.unlink(self._tmp.name)
        self.nx, self.ny = nx, ny
        self.stack_mb = max(64, (nx * ny * 8 * 8) // (1 << 20) + 64)
        n = nx * ny
        shape = (ny, nx)

        def plane(sym, ct):
            return np.ctypeslib.as_array((ct * n).in_dll(self.L, sym)).reshape(shape)
        self.u, self.v = plane("g_u", C.c_float), plane("g_v", C.c_float)
        self.utmp, self.vtmp = plane("g_utmp", C.c_float), plane("g_vtmp", C.c_float)
        self.solid, self.source, self.sink = (plane(s, C.c_uint8) for s in ("g_solid", "g_source", "g_sink"))
        self.count, self.prev_count = plane("g_marker_count", C.c_uint8), plane("g_prev_marker_count", C.c_uint8)
        self.precon, self.q = plane("g_precon", C.c_double), plane("g_q", C.c_double)
        self.adiag = plane("g_a", C.c_int8)
        self.cr, self.cg, self.cb = (plane(s, C.c_float) for s in ("g_r", "g_g", "g_b"))
        self._rainbow = C.c_bool.in_dll(self.L, "g_rainbow_enabled")
        self._frame_count = C.c_uint16.in_dll(self.L, "g_frame_count")
        self._markers = np.ctypeslib.as_array((C.c_float * (8 * n)).in_dll(self.L, "g_markers")).reshape(-1, 2)
        self._len = C.c_size_t.in_dll(self.L, "g_markers_length")
        self._exhausted = C.c_bool.in_dll(self.L, "g_source_exhausted")
        # randf()'s function-local static: located through the symbol offsets that
        # build_ref.sh recorded next to the library
        self._rng = None
        syms = {}
        with open(path[:-3] + ".syms") as f:
            for line in f:
                a, _, name = line.split()
                syms[name.split(".")[0]] = int(a, 16)
        if "rng_state" in syms and "g_u" in syms:
            base = C.addressof((C.c_float * n).in_dll(self.L, "g_u")) - syms["g_u"]
            self._rng = C.c_uint64.from_address(base + syms["rng_state"])
        L = self.L
        L.sim_init.argtypes = [_Args]
        L.calculate_timestep.restype = C.c_float; L.calculate_timestep.argtypes = [C.c_float]
        L.advect_markers.argtypes = [C.c_float]
        FP = C.c_void_p
        L.extrapolate.argtypes = [FP, C.c_int]; L.zero_bounds.argtypes = [FP, C.c_int]
        L.advect_u.argtypes = [FP, FP, C.c_float, FP]; L.advect_v.argtypes = [FP, FP, C.c_float, FP]
        L.apply_body_forces.argtypes = [FP, C.c_float]
        L.project.argtypes = [C.c_float, FP, FP, FP, FP]
        L.apply_preconditioner.argtypes = [FP, FP]; L.apply_a.argtypes = [FP, FP]
        L.dot.restype = C.c_double; L.dot.argtypes = [FP, FP]
        L.interpolate.restype = C.c_float

    @property
    def n_markers(self): return int(self._len.value)
    @property
    def markers(self): return self._markers[: self.n_markers]

    def set_markers(self, m):
        m = np.ascontiguousarray(m, dtype=np.float32).reshape(-1, 2)
        self._markers[: len(m)] = m
        self._len.value = len(m)

    @property
    def rng_state(self): return int(self._rng.value)
    @rng_state.setter
    def rng_state(self, v): self._rng.value = v
    @property
    def source_exhausted(self): return bool(self._exhausted.value)

    def init_from_text(self, text, rainbow=False):
        import tempfile
        if isinstance(text, str):
            text = text.encode()
        self._rainbow.value = bool(rainbow)          # main() sets the global before sim_init (main.c:1020)
        with tempfile.NamedTemporaryFile(suffix=".txt", delete=False) as f:
            f.write(text)
            name = f.name
        try:
            a = _Args(name.encode(), bool(rainbow))
            call_with_big_stack(self.L.sim_init, a, stack_mb=self.stack_mb)
        finally:
            os.unlink(name)

    def _p(self, a): return a.ctypes.data

    def calculate_timestep(self, frame_time=0.1): return float(self.L.calculate_timestep(np.float32(frame_time)))
    def advect_markers(self, dt): self.L.advect_markers(np.float32(dt))
    def refresh_marker_counts(self): self.L.refresh_marker_counts()
    def update_fluid_sources(self): self.L.update_fluid_sources()
    def extrapolate(self, q, t): self.L.extrapolate(self._p(q), t)
    def zero_bounds(self, q, t): self.L.zero_bounds(self._p(q), t)
    def advect_u(self, dt): self.L.advect_u(self._p(self.u), self._p(self.v), np.float32(dt), self._p(self.utmp))
    def advect_v(self, dt): self.L.advect_v(self._p(self.u), self._p(self.v), np.float32(dt), self._p(self.vtmp))
    def apply_body_forces(self, dt): self.L.apply_body_forces(self._p(self.vtmp), np.float32(dt))

    def project(self, dt):
        call_with_big_stack(self.L.project, np.float32(dt), self._p(self.utmp), self._p(self.vtmp),
                            self._p(self.u), self._p(self.v), stack_mb=self.stack_mb)

    def apply_preconditioner(self, r, z): self.L.apply_preconditioner(self._p(r), self._p(z))
    def apply_a(self, s, out): self.L.apply_a(self._p(s), self._p(out))
    def dot(self, a, b): return float(self.L.dot(self._p(a), self._p(b)))
    def step_frame(self): call_with_big_stack(self.L.sim_step, stack_mb=self.stack_mb)
