"""The host renderer (euler_b200/host/render.c) against the reference's own draw code
(reposition_cursor + draw_rows + hide_cursor, main.c:914-959), byte for byte: same escape
sequences, same clipping to the terminal window, same behaviour around sinks, and with
--rainbow the same 24-bit colours (main.c:902-912, misc/color.h).  The reference runs in-process
from oracle/_ref (unmodified main.c); rendering is host code, off the timed path."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT, SCENARIOS
from euler_b200 import shipped_text
from oracle.oracle import Reference, ref_available

pytestmark = pytest.mark.skipif(not ref_available(100, 40), reason="oracle/_ref not built (no /root/reference)")


class _Buf(C.Structure):                       # buffer_t, misc/terminal.h:3-6
    _fields_ = [("data", C.c_void_p), ("len", C.c_int)]


class _Screen(C.Structure):                    # euler_screen, render.h
    _fields_ = [("buf", C.c_void_p), ("len", C.c_size_t), ("cap", C.c_size_t), ("cols", C.c_int), ("rows", C.c_int)]


@pytest.fixture(scope="module")
def host():
    L = C.CDLL(os.path.join(ROOT, "euler_b200", "lib", "libeuler_host.so"))
    vp = C.c_void_p
    L.euler_draw.argtypes = [C.POINTER(_Screen), C.c_int, C.c_int, vp, vp, vp]
    L.euler_draw_rainbow.argtypes = [C.POINTER(_Screen), C.c_int, C.c_int, vp, vp, vp, vp, vp, vp]
    L.euler_screen_free.argtypes = [C.POINTER(_Screen)]
    L.euler_color_byte.argtypes = [C.c_float]
    return L


def _reference_frame(r, cols, rows):
    C.c_int.in_dll(r.L, "g_wx").value = cols
    C.c_int.in_dll(r.L, "g_wy").value = rows
    buf = _Buf(None, 0)
    for fn in (r.L.reposition_cursor, r.L.draw_rows, r.L.hide_cursor):
        fn.argtypes = [C.POINTER(_Buf)]
        fn(C.byref(buf))
    out = C.string_at(buf.data, buf.len)
    r.L.buffer_free.argtypes = [C.POINTER(_Buf)]
    r.L.buffer_free(C.byref(buf))
    return out


def _our_frame(host, r, cols, rows, rainbow):
    scr = _Screen(None, 0, 0, cols, rows)
    planes = [np.ascontiguousarray(a) for a in (r.solid, r.sink, r.count)]
    args = [C.byref(scr), 100, 40] + [a.ctypes.data for a in planes]
    if rainbow:
        colour = [np.ascontiguousarray(a) for a in (r.cr, r.cg, r.cb)]
        host.euler_draw_rainbow(*(args + [a.ctypes.data for a in colour]))
    else:
        host.euler_draw(*args)
    out = C.string_at(scr.buf, scr.len)
    host.euler_screen_free(C.byref(scr))
    return out


@pytest.mark.parametrize("rainbow", [False, True])
@pytest.mark.parametrize("name", SCENARIOS)
def test_frames_are_the_references_bytes(host, name, rainbow, capfd):
    r = Reference(100, 40)
    r.init_from_text(shipped_text(name), rainbow=rainbow)
    windows = [(200, 60), (98, 38), (80, 24), (40, 10), (1, 1), (0, 0), (98, 39), (99, 37)]
    for frame in range(0, 36):
        if frame in (0, 1, 5, 20, 35):
            for cols, rows in windows:
                want = _reference_frame(r, cols, rows)
                got = _our_frame(host, r, cols, rows, rainbow)
                assert got == want, (name, frame, cols, rows)
        r.step_frame()
    capfd.readouterr()                         # the frames our renderer wrote to stdout


def test_water_after_a_sink_and_colour_bytes(host, capfd):
    """The quirk the byte comparison pins (main.c:928-932: a sink resets the colour but not
    prev_water) on a hand-made row, and euler_color_byte == float_to_byte_color(linear_to_sRGB)."""
    r = Reference(100, 40)
    r.init_from_text("00=00X0 0=\n")
    assert r.count[38, 1] and r.sink[38, 3] and r.count[38, 4]
    want = _reference_frame(r, 98, 38)
    assert want.startswith(b"\x1b[H\x1b[34m00\x1b[0m=00\x1b[0mX\x1b[34m0\x1b[0m \x1b[34m0\x1b[0m=")
    assert _our_frame(host, r, 98, 38, False) == want
    capfd.readouterr()
    end = np.nextafter(np.float32(256), np.float32(0))
    for x in (0.0, 1.0, 2.0, 0.5, 0.2, 1e-6, 0.999):           # misc/color.h:6-13 in fp32
        want = int(min(max(end * np.power(np.float32(x), np.float32(1 / 2.2)), np.float32(0)), end))
        assert host.euler_color_byte(x) == want, x
    assert host.euler_color_byte(0.0) == 0 and host.euler_color_byte(1.0) == 255 and host.euler_color_byte(7.0) == 255
