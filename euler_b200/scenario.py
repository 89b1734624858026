"""ctypes mirror of euler_b200/host/scenario.{h,c} (the host C scenario parser / marker
seeding, reference sim_init main.c:209-274).  No logic of its own: one source of truth in C."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SCENARIO_DIR = os.path.join(os.path.dirname(_HERE), "tests", "golden")
_LIB = None


class _Scn(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int),
                ("solid", C.POINTER(C.c_uint8)), ("source", C.POINTER(C.c_uint8)),
                ("sink", C.POINTER(C.c_uint8)), ("fluid", C.POINTER(C.c_uint8)),
                ("markers", C.POINTER(C.c_float)), ("n_markers", C.c_size_t),
                ("rng_state", C.c_uint64)]


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "lib", "libeuler_host.so")
        if not os.path.exists(path):
            raise RuntimeError("%s missing: run `make host`" % path)
        L = C.CDLL(path)
        L.euler_scenario_from_text.argtypes = [C.POINTER(_Scn), C.c_char_p, C.c_long, C.c_int, C.c_int]
        L.euler_scenario_free.argtypes = [C.POINTER(_Scn)]
        L.euler_scenario_markers_row_major.argtypes = [C.POINTER(_Scn)]
        L.euler_scenario_resample.restype = C.c_void_p
        L.euler_scenario_resample.argtypes = [C.c_char_p, C.c_long, C.c_int, C.c_int, C.POINTER(C.c_long)]
        L.euler_scenario_synthetic.restype = C.c_void_p
        L.euler_scenario_synthetic.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_long)]
        L.euler_scenario_export.restype = C.c_void_p
        L.euler_scenario_export.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.POINTER(C.c_long)]
        L.euler_randf.restype = C.c_float
        L.euler_randf.argtypes = [C.POINTER(C.c_uint64)]
        _LIB = L
    return _LIB


_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


def _take(ptr, n):
    try:
        return C.string_at(ptr, n)
    finally:
        _libc.free(ptr)


def resample(text, out_w, out_h):
    """Nearest-neighbour resample of a scenario text to out_w x out_h characters."""
    if isinstance(text, str):
        text = text.encode()
    n = C.c_long(0)
    p = _lib().euler_scenario_resample(text, len(text), out_w, out_h, C.byref(n))
    if not p:
        raise ValueError("empty scenario")
    return _take(p, n.value)


def synthetic(name, nx, ny):
    n = C.c_long(0)
    p = _lib().euler_scenario_synthetic(name.encode(), nx, ny, C.byref(n))
    if not p:
        raise ValueError("unknown synthetic scenario %r" % name)
    return _take(p, n.value)


def export_text(solid, source, sink, count):
    """The current state in the scenario-file format (euler_scenario_export): static masks plus
    '0' for every cell that holds markers.  Arrays are [ny][nx] uint8."""
    planes = [np.ascontiguousarray(a, dtype=np.uint8) for a in (solid, source, sink, count)]
    ny, nx = planes[0].shape
    if any(a.shape != (ny, nx) for a in planes):
        raise ValueError("planes must have the same [ny][nx] shape")
    n = C.c_long(0)
    p = _lib().euler_scenario_export(nx, ny, *(a.ctypes.data for a in planes), C.byref(n))
    if not p:
        raise ValueError("grid too small to export")
    return _take(p, n.value)


def shipped_text(name):
    """Text of one of the five scenario files the reference ships (committed as an input
    fixture in tests/golden/scenarios.json by tests/golden/make_golden.py)."""
    import json
    with open(os.path.join(SCENARIO_DIR, "scenarios.json")) as f:
        return json.load(f)[name].encode("ascii")


class Scenario:
    """Parsed scenario: static masks, seeded markers and the RNG state after seeding."""

    def __init__(self, text, nx, ny, row_major_markers=False):
        if isinstance(text, str):
            text = text.encode()
        s = _Scn()
        rc = _lib().euler_scenario_from_text(C.byref(s), text, len(text), nx, ny)
        if rc:
            raise MemoryError("scenario_from_text failed (%d)" % rc)
        if row_major_markers and _lib().euler_scenario_markers_row_major(C.byref(s)):
            _lib().euler_scenario_free(C.byref(s))
            raise MemoryError("markers_row_major failed")
        try:
            shape = (ny, nx)
            self.nx, self.ny = nx, ny
            self.solid = np.ctypeslib.as_array(s.solid, shape=shape).copy()
            self.source = np.ctypeslib.as_array(s.source, shape=shape).copy()
            self.sink = np.ctypeslib.as_array(s.sink, shape=shape).copy()
            self.fluid = np.ctypeslib.as_array(s.fluid, shape=shape).copy()
            n = int(s.n_markers)
            self.markers = (np.ctypeslib.as_array(s.markers, shape=(max(n, 1), 2))[:n]).copy()
            self.rng_state = int(s.rng_state)
        finally:
            _lib().euler_scenario_free(C.byref(s))

    @classmethod
    def from_file(cls, path, nx, ny):
        with open(path, "rb") as f:
            return cls(f.read(), nx, ny)

    @classmethod
    def shipped(cls, name, nx=100, ny=40):
        """One of the scenario files shipped with the reference, optionally resampled."""
        text = shipped_text(name)
        if (nx, ny) != (100, 40):
            text = resample(text, nx - 2, ny - 2)
        return cls(text, nx, ny)
