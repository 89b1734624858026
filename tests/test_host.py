"""Host C side (euler_b200/host/scenario.c): parser, ring of sinks, marker seeding, RNG and the
resampler, against the oracle's restatement of sim_init (main.c:209-274)."""
import numpy as np
import pytest

from conftest import SCENARIOS, same_bits
from euler_b200 import Scenario, shipped_text, resample, synthetic
from oracle.oracle import Oracle


@pytest.mark.parametrize("name", SCENARIOS)
def test_parser_and_seeding_match_oracle(name):
    text = shipped_text(name)
    s, o = Scenario(text, 100, 40), Oracle(100, 40, text)
    assert np.array_equal(s.solid, o.solid) and np.array_equal(s.source, o.source)
    assert np.array_equal(s.sink, o.sink)
    # the oracle already ran refresh_marker_counts, which deletes nothing at init
    assert same_bits(s.markers, o.markers)
    assert s.rng_state == int(o.c.rng_state)
    assert s.sink[0].all() and s.sink[-1].all() and s.sink[:, 0].all() and s.sink[:, -1].all()


def test_other_grid_sizes_and_truncation():
    text = shipped_text("block")          # 95 columns of text into a 64-wide grid: truncated
    s, o = Scenario(text, 64, 48), Oracle(64, 48, text)
    assert np.array_equal(s.solid, o.solid) and same_bits(s.markers, o.markers)
    s, o = Scenario(text, 130, 20), Oracle(130, 20, text)   # wider than the text, fewer rows
    assert np.array_equal(s.solid, o.solid) and same_bits(s.markers, o.markers)


def test_empty_and_ragged_input():
    s = Scenario("", 16, 16)
    assert len(s.markers) == 0 and not s.solid.any() and s.sink.sum() == 16 * 4 - 4
    s = Scenario("X0\n\n?\n   =\n", 16, 16)
    o = Oracle(16, 16, "X0\n\n?\n   =\n")
    assert np.array_equal(s.solid, o.solid) and np.array_equal(s.source, o.source)
    assert np.array_equal(s.sink, o.sink) and same_bits(s.markers, o.markers)
    assert len(s.markers) == 8


def test_resample_identity_and_shape():
    text = shipped_text("basic")
    lines = text.decode().split("\n")
    h = len([l for l in lines if l != ""])
    w = max(len(l) for l in lines)
    same = resample(text, w, h).decode().split("\n")
    assert [l.ljust(w) for l in lines if l != ""] == same[:h]
    big = resample(text, 3 * w, 2 * h).decode().split("\n")
    assert len(big[0]) == 3 * w and len([l for l in big if l]) == 2 * h
    assert big[0] == "X" * (3 * w)


def test_synthetic_scenarios():
    t = synthetic("basic-fill", 64, 64).decode()
    rows = t.strip("\n").split("\n")
    assert len(rows) == 62 and all(len(r) == 62 for r in rows)
    assert rows[0] == "X" * 62 and rows[-1] == "X" * 62
    s = Scenario(t, 64, 64)
    assert 0.15 < s.fluid.mean() < 0.25
    with pytest.raises(ValueError):
        synthetic("nope", 8, 8)


def test_row_major_marker_order_is_a_permutation():
    text = shipped_text("weird-edges")
    a, b = Scenario(text, 100, 40), Scenario(text, 100, 40, row_major_markers=True)
    assert a.rng_state == b.rng_state and len(a.markers) == len(b.markers)
    key = lambda m: m[np.lexsort((m[:, 0], m[:, 1]))]
    assert same_bits(key(a.markers), key(b.markers))
    cells = (np.floor(b.markers[:, 1]).astype(np.int64) * 100 + np.floor(b.markers[:, 0]).astype(np.int64))
    assert (np.diff(cells) >= 0).all()          # row-major, 4 per cell


def test_parser_fuzz_against_oracle():
    """Random scenario texts over the format's alphabet plus junk, ragged and over-long lines,
    empty lines, no trailing newline; random grid sizes: masks, seeded markers and the RNG state
    equal the oracle's restatement of sim_init (main.c:209-274)."""
    from hypothesis import given, settings, strategies as st

    line = st.text(alphabet="X0?= ab\t", min_size=0, max_size=40)

    @settings(max_examples=150, deadline=None)
    @given(st.lists(line, min_size=0, max_size=30), st.booleans(), st.integers(4, 48), st.integers(4, 36))
    def run(lines, trailing, nx, ny):
        text = "\n".join(lines) + ("\n" if trailing else "")
        s, o = Scenario(text, nx, ny), Oracle(nx, ny, text)
        assert np.array_equal(s.solid, o.solid) and np.array_equal(s.source, o.source)
        assert np.array_equal(s.sink, o.sink)
        assert same_bits(s.markers, o.markers) and s.rng_state == int(o.c.rng_state)

    run()


@pytest.mark.parametrize("name", SCENARIOS)
def test_scenario_export_round_trip(name):
    """euler_scenario_export is the inverse of the parser: exporting the initial
    state and parsing it again reproduces masks, fluid cells, seeded markers and RNG state;
    exporting a LATER state gives a scenario whose fluid cells are the cells holding markers."""
    from euler_b200 import export_text
    text = shipped_text(name)
    s0 = Scenario(text, 100, 40)
    out = export_text(s0.solid, s0.source, s0.sink, s0.fluid)
    rows = out.decode().split("\n")
    assert len(rows) == 39 and rows[-1] == "" and all(len(r) == 98 for r in rows[:-1])
    assert set(out.decode()) <= set("X0?= \n")
    s1 = Scenario(out, 100, 40)
    for f in ("solid", "source", "sink", "fluid"):
        assert np.array_equal(getattr(s0, f), getattr(s1, f)), f
    assert same_bits(s0.markers, s1.markers) and s0.rng_state == s1.rng_state
    assert export_text(s1.solid, s1.source, s1.sink, s1.fluid) == out          # idempotent
    # a later state (from the oracle): fluid of the re-parsed export == cells with markers
    o = Oracle(100, 40, text)
    for _ in range(25):
        o.step_frame()
    s2 = Scenario(export_text(o.solid, o.source, o.sink, o.count), 100, 40)
    assert np.array_equal(s2.solid, o.solid) and np.array_equal(s2.source, o.source) and np.array_equal(s2.sink, o.sink)
    expect = ((o.count != 0) & (o.solid == 0) & (o.sink == 0)) | (o.source != 0)
    assert np.array_equal(s2.fluid != 0, expect)
    assert len(s2.markers) == 4 * int(expect.sum())


def test_scenario_export_other_sizes_and_errors():
    from euler_b200 import export_text
    text = resample(shipped_text("weird-edges"), 62, 46)
    s0 = Scenario(text, 64, 48)
    s1 = Scenario(export_text(s0.solid, s0.source, s0.sink, s0.fluid), 64, 48)
    assert np.array_equal(s0.solid, s1.solid) and np.array_equal(s0.fluid, s1.fluid) and same_bits(s0.markers, s1.markers)
    z = np.zeros((2, 2), np.uint8)
    with pytest.raises(ValueError):
        export_text(z, z, z, z)
    with pytest.raises(ValueError):
        export_text(np.zeros((5, 5), np.uint8), z, z, z)


def test_file_rendezvous(tmp_path):
    """euler_b200/host/rendezvous.c: the side channel of the multi-process host runs — atomic
    publish, blocking fetch, size check, timeout."""
    import ctypes as C
    import os
    import threading
    import time
    from euler_b200.scenario import _lib
    L = _lib()
    L.euler_rdv_publish.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_void_p, C.c_size_t]
    L.euler_rdv_fetch.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_void_p, C.c_size_t, C.c_int]
    d = str(tmp_path).encode()
    payload = bytes(range(256)) * 3
    assert L.euler_rdv_publish(d, b"blob", 3, payload, len(payload)) == 0
    assert sorted(os.listdir(tmp_path)) == ["blob3.bin"]                      # no temporary left behind
    buf = C.create_string_buffer(len(payload))
    assert L.euler_rdv_fetch(d, b"blob", 3, buf, len(payload), 1) == 0 and buf.raw == payload
    assert L.euler_rdv_fetch(d, b"blob", 3, buf, len(payload) - 1, 1) == -1   # size mismatch
    t0 = time.time()
    assert L.euler_rdv_fetch(d, b"blob", 4, buf, 8, 1) == -2                  # nobody publishes: timeout
    assert 0.9 <= time.time() - t0 < 3.0
    assert L.euler_rdv_publish(str(tmp_path / "missing").encode(), b"x", 0, payload, 4) == -1
    # a fetch that starts before the publish
    def late():
        time.sleep(0.3)
        L.euler_rdv_publish(d, b"uid", 0, payload, 128)
    th = threading.Thread(target=late)
    th.start()
    got = C.create_string_buffer(128)
    assert L.euler_rdv_fetch(d, b"uid", 0, got, 128, 10) == 0 and got.raw == payload[:128]
    th.join()
    assert L.euler_rdv_publish(d, b"empty", 1, None, 0) == 0 and L.euler_rdv_fetch(d, b"empty", 1, None, 0, 1) == 0


def _handshake_proc(d, rank, ranks, uid, delay, out):
    import ctypes as C
    import time
    from euler_b200.scenario import _lib
    L = _lib()
    L.euler_rdv_handshake.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64), C.c_int]
    time.sleep(delay)
    buf = C.create_string_buffer(uid if rank == 0 else b"\0" * 128, 128)
    key = C.c_uint64(0)
    rc = L.euler_rdv_handshake(d.encode(), rank, ranks, buf, 128, C.byref(key), 20)
    out.put((rank, rc, buf.raw, key.value))


def test_rendezvous_handshake_ignores_a_stale_directory(tmp_path):
    """Two runs, one after the other, in the SAME directory (a reused DIR let ranks
    pick up the previous run's communicator id and hang in ncclCommInitRank): in the second run the
    non-zero ranks start BEFORE rank 0, with the first run's answer / hello / ack files still there,
    and must end up with the second run's id and key; keyed files of the first run are ignored."""
    import ctypes as C
    import multiprocessing as mp
    from euler_b200.scenario import _lib
    L = _lib()
    L.euler_rdv_publish_keyed.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_uint64, C.c_void_p, C.c_size_t]
    L.euler_rdv_fetch_keyed.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_uint64, C.c_void_p, C.c_size_t, C.c_int]
    d = str(tmp_path)
    ctx = mp.get_context("spawn")
    keys = []
    for run, delays in enumerate([(0.0, 0.1, 0.2), (0.6, 0.0, 0.0)]):
        uid = bytes([run + 1]) * 128
        out = ctx.Queue()
        procs = [ctx.Process(target=_handshake_proc, args=(d, r, 3, uid, delays[r], out)) for r in range(3)]
        for p in procs:
            p.start()
        res = sorted(out.get(timeout=60) for _ in procs)
        for p in procs:
            p.join()
        assert [x[1] for x in res] == [0, 0, 0]
        assert all(x[2] == uid for x in res), "a rank accepted another run's communicator id"
        assert len({x[3] for x in res}) == 1
        keys.append(res[0][3])
        payload = bytes([0x40 + run]) * 16
        assert L.euler_rdv_publish_keyed(d.encode(), b"info", 1, keys[-1], payload, 16) == 0
    assert keys[0] != keys[1]
    buf = C.create_string_buffer(16)
    assert L.euler_rdv_fetch_keyed(d.encode(), b"info", 1, keys[1], buf, 16, 1) == 0 and buf.raw == b"\x41" * 16
    assert L.euler_rdv_fetch_keyed(d.encode(), b"info", 1, keys[0], buf, 16, 1) == -2     # first run's file is gone / not ours
    # a rank that never shows up: timeout, not a hang
    L.euler_rdv_handshake.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64), C.c_int]
    key = C.c_uint64(0)
    lone = C.create_string_buffer(b"\x07" * 128, 128)
    assert L.euler_rdv_handshake(str(tmp_path / ".").encode(), 0, 4, lone, 128, C.byref(key), 1) == -2


def test_host_program_rank_arguments():
    """bin/euler-gpu refuses an incomplete --ranks set-up before it touches a GPU."""
    import os
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "bin", "euler-gpu")
    assert os.path.exists(exe), "run `make host`"
    for args in (["--ranks", "2", "--rank", "0", "x"], ["--headless", "--ranks", "2", "--rank", "2", "--rendezvous", "/tmp", "x"],
                 ["--headless", "--ranks", "2", "--rank", "0", "x"]):
        r = subprocess.run([exe] + args, capture_output=True, text=True, cwd=ROOT)
        assert r.returncode == 1 and "--ranks needs" in r.stderr, args


def test_host_program_cli_errors_like_the_reference():
    """parse_args / sim_init error behaviour (main.c:982-999, 213-214): usage line and exit 1
    without arguments, "Unrecognized input: X" for an unknown option, "Could not load F!" for a
    missing scenario — all before any GPU work."""
    import os
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "bin", "euler-gpu")
    run = lambda *a: subprocess.run([exe] + list(a), capture_output=True, text=True, cwd=ROOT)
    r = run()
    assert r.returncode == 1 and r.stderr.startswith("usage: %s [--rainbow]" % exe) and "<scenario>" in r.stderr
    r = run("--bogus", "x.txt")
    assert r.returncode == 1 and r.stderr == "Unrecognized input: --bogus\n"
    r = run("--headless", "/nonexistent/scenario.txt")
    assert r.returncode == 1 and r.stderr == "Could not load /nonexistent/scenario.txt!\n"
    r = run("--precon", "nope", "x.txt")
    assert r.returncode == 1 and r.stderr.startswith("usage:")
