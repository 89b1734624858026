"""euler_b200/host/checkpoint.c on the CPU: the two calls only need six entry points of the
C-ABI (euler_ckpt_api), so a ctypes fake of get/set/stats stands in for libeuler_gpu.so.
Round trip, the refusals, and handles without the fp64 g_precon plane (pcg_dtype = FP32)."""
import ctypes as C
import os

import numpy as np
import pytest

from euler_b200 import gpu as G        # Stats struct + field ids (loads the library, no device needed)

HOST_LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        "euler_b200", "lib", "libeuler_host.so")
GET = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t)
SET = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t)
STATS = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(G.Stats))
SET64 = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64)
SETI = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int)


class Api(C.Structure):
    _fields_ = [("get", GET), ("set", SET), ("stats", STATS), ("set_rng_state", SET64),
                ("set_source_exhausted", SETI), ("set_frame_count", SET64)]


class FakeSim:
    """In-memory stand-in for a handle: planes by field id, markers, the scalars."""
    DT = {G.F_U: np.float32, G.F_V: np.float32, G.F_COUNT: np.uint8, G.F_PREV_COUNT: np.uint8,
          G.F_PRECON: np.float64, G.F_CR: np.float32, G.F_CG: np.float32, G.F_CB: np.float32}

    def __init__(self, nx, ny, seed, rainbow=False, precon=True, n_markers=37, mask_seed=0):
        rng = np.random.default_rng(seed)
        mrng = np.random.default_rng(1000 + mask_seed)      # the static masks = "the scenario"
        self.masks = {f: (mrng.random((ny, nx)) < 0.2).astype(np.uint8) for f in (G.F_SOLID, G.F_SOURCE, G.F_SINK)}
        fields = [G.F_U, G.F_V, G.F_COUNT, G.F_PREV_COUNT] + ([G.F_PRECON] if precon else []) + \
                 ([G.F_CR, G.F_CG, G.F_CB] if rainbow else [])
        self.planes = {}
        for f in fields:
            dt = self.DT[f]
            self.planes[f] = (rng.integers(0, 5, (ny, nx)).astype(dt) if dt == np.uint8
                              else rng.standard_normal((ny, nx)).astype(dt))
        self.markers = rng.random((n_markers, 2)).astype(np.float32)
        self.rng_state, self.frames, self.exhausted = int(rng.integers(1, 2 ** 62)), int(rng.integers(0, 999)), int(seed & 1)
        self.api = Api(GET(self._get), SET(self._set), STATS(self._stats), SET64(self._rng),
                       SETI(self._exh), SET64(self._frames))

    def _get(self, _h, field, dst, n):
        if field == G.F_MARKERS:
            src = self.markers
        elif field in self.planes:
            src = self.planes[field]
        elif field in self.masks:
            src = self.masks[field]
        else:
            return -1                                   # EULER_E_INVALID: no such plane on this handle
        if n != src.nbytes:
            return -1
        C.memmove(dst, src.ctypes.data, n)
        return 0

    def _set(self, _h, field, src, n):
        if field == G.F_MARKERS:
            self.markers = np.frombuffer(C.string_at(src, n), dtype=np.float32).reshape(-1, 2).copy()
            return 0
        if field not in self.planes or n != self.planes[field].nbytes:
            return -1
        C.memmove(self.planes[field].ctypes.data, src, n)
        return 0

    def _stats(self, _h, out):
        out.contents.n_markers = len(self.markers)
        out.contents.rng_state = self.rng_state
        out.contents.frames = self.frames
        out.contents.source_exhausted = self.exhausted
        return 0

    def _rng(self, _h, v): self.rng_state = v; return 0
    def _exh(self, _h, v): self.exhausted = v; return 0
    def _frames(self, _h, v): self.frames = v; return 0


@pytest.fixture(scope="module")
def lib():
    L = C.CDLL(HOST_LIB)
    for fn in (L.euler_checkpoint_save, L.euler_checkpoint_load):
        fn.argtypes = [C.POINTER(Api), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_char_p]
        fn.restype = C.c_int
    return L


def _same(a, b):
    return (all(np.array_equal(a.planes[f], b.planes[f]) for f in a.planes if f in b.planes)
            and np.array_equal(a.markers, b.markers)
            and (a.rng_state, a.frames, a.exhausted) == (b.rng_state, b.frames, b.exhausted))


@pytest.mark.parametrize("rainbow", [False, True])
def test_round_trip(lib, tmp_path, rainbow):
    nx, ny = 23, 11
    a, b = FakeSim(nx, ny, 1, rainbow), FakeSim(nx, ny, 2, rainbow, n_markers=5)
    path = str(tmp_path / "state.ck").encode()
    assert lib.euler_checkpoint_save(C.byref(a.api), None, nx, ny, int(rainbow), path) == 0
    assert not _same(a, b)
    assert lib.euler_checkpoint_load(C.byref(b.api), None, nx, ny, int(rainbow), path) == 0
    assert _same(a, b) and set(a.planes) == set(b.planes)


def test_refusals(lib, tmp_path):
    nx, ny = 16, 8
    a = FakeSim(nx, ny, 3)
    path = str(tmp_path / "s.ck").encode()
    assert lib.euler_checkpoint_save(C.byref(a.api), None, nx, ny, 0, path) == 0
    assert lib.euler_checkpoint_load(C.byref(FakeSim(nx + 1, ny, 4).api), None, nx + 1, ny, 0, path) == -2
    assert lib.euler_checkpoint_load(C.byref(FakeSim(nx, ny, 4, True).api), None, nx, ny, 1, path) == -2
    assert lib.euler_checkpoint_load(C.byref(a.api), None, nx, ny, 0, str(tmp_path / "missing").encode()) == -1
    bad = tmp_path / "bad.ck"
    bad.write_bytes(b"NOTEULER" + b"\0" * 64)
    assert lib.euler_checkpoint_load(C.byref(a.api), None, nx, ny, 0, str(bad).encode()) == -1
    # same size, another scenario (other static masks): the dynamic state does not belong there
    assert lib.euler_checkpoint_load(C.byref(FakeSim(nx, ny, 4, mask_seed=1).api), None, nx, ny, 0, path) == -2
    cut = tmp_path / "cut.ck"
    cut.write_bytes(open(path.decode(), "rb").read()[:200])       # truncated planes
    assert lib.euler_checkpoint_load(C.byref(FakeSim(nx, ny, 5).api), None, nx, ny, 0, str(cut).encode()) == -1


def test_handles_without_the_fp64_precon_plane(lib, tmp_path):
    """pcg_dtype = FP32 handles have no g_precon plane: save writes zeros and flags it; such a
    file loads into either kind of handle, and a file WITH the plane loads into a handle
    without it (the plane is skipped, everything else restored)."""
    nx, ny = 20, 9
    mixed, full = FakeSim(nx, ny, 6, precon=False), FakeSim(nx, ny, 7)
    p_mixed, p_full = str(tmp_path / "m.ck").encode(), str(tmp_path / "f.ck").encode()
    assert lib.euler_checkpoint_save(C.byref(mixed.api), None, nx, ny, 0, p_mixed) == 0
    assert lib.euler_checkpoint_save(C.byref(full.api), None, nx, ny, 0, p_full) == 0
    assert os.path.getsize(p_mixed) == os.path.getsize(p_full)       # fixed layout (same marker count)
    # mixed file -> fp64 handle: the handle's g_precon restarts from zero like a fresh run's
    # (main.c:577) instead of keeping whatever it held (the IC(0) iterates depend on it)
    tgt = FakeSim(nx, ny, 8)
    assert lib.euler_checkpoint_load(C.byref(tgt.api), None, nx, ny, 0, p_mixed) == 0
    assert _same(mixed, tgt) and not tgt.planes[G.F_PRECON].any()
    # fp64 file -> mixed handle
    tgt = FakeSim(nx, ny, 9, precon=False)
    assert lib.euler_checkpoint_load(C.byref(tgt.api), None, nx, ny, 0, p_full) == 0
    assert _same(full, tgt) and G.F_PRECON not in tgt.planes
