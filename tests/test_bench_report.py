"""bench.py's roofline arithmetic on a recorded profile (no GPU): the per-kernel figures of a
committed bench line are reproduced from its own launch times and unit counts."""
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.load(f)


def test_report_reproduces_the_committed_fp64_line():
    b = _bench()
    d = _line("r01m_bench_16384_rb.json")
    prof = {k: (v["ms_avg"] * v["launches"], v["launches"]) for k, v in d["kernels"].items()}
    cfg = d["config"]
    kernels, roof = b.roofline_report(prof, b.ALG_BYTES_PER_CELL, cfg["active_cells"], 16384 * 16384,
                                      cfg["markers"], headline=True, mixed=False)
    assert roof["kernel"] == d["roofline"]["kernel"] == "axpy_norm"
    assert roof["bytes_per_unit"] == 40.0 and roof["units_per_launch"] == cfg["active_cells"]
    for name, rec in d["kernels"].items():
        if rec["gbs"] is None:
            assert kernels[name]["gbs"] is None
            continue
        if name in b.DRY_BYTES_PER_CELL:
            # grid stages: `frac` is what the kernel has to move (full bytes on wet cells, mask + zero
            # store elsewhere: the round-1 line's gbs_touched); the dense SURVEY 8d figure that line
            # reported as its frac is kept as frac_dense
            assert abs(kernels[name]["gbs"] - rec["gbs_touched"]) <= 2e-3 * rec["gbs_touched"], name
            assert abs(kernels[name]["frac_dense"] - rec["frac"]) <= 2e-3, name
            assert kernels[name]["frac"] < 1.0
        else:
            assert abs(kernels[name]["gbs"] - rec["gbs"]) <= 2e-3 * rec["gbs"], name   # ms_avg is rounded in the file
        assert abs(kernels[name]["frac_nominal"] - kernels[name]["gbs"] / 8000.0) < 1e-3
    assert roof["traffic"] == 2168300000.0 and roof["peak_nominal"] == 8000.0
    assert abs(sum(k["share"] for k in kernels.values()) - 1.0) < 1e-2


def test_report_uses_the_fp32_byte_counts_and_traffic_in_mixed_mode():
    b = _bench()
    d = _line("r01n_bench_16384_rb_fp32.json")
    prof = {k: (v["ms_avg"] * v["launches"], v["launches"]) for k, v in d["kernels"].items()}
    cfg = d["config"]
    alg = dict(b.ALG_BYTES_PER_CELL, **b.ALG_BYTES_PER_CELL_FP32)
    kernels, roof = b.roofline_report(prof, alg, cfg["active_cells"], 16384 * 16384, cfg["markers"],
                                      headline=True, mixed=True)
    assert roof["kernel"] == "axpy_norm" and roof["bytes_per_unit"] == 25.0
    assert roof["traffic"] == 1306900000.0
    assert abs(kernels["rb_forward"]["gbs"] - 13.0 * cfg["active_cells"] / d["kernels"]["rb_forward"]["ms_avg"] / 1e6) < 1.0
    assert "true_residual" in kernels and kernels["true_residual"]["gbs"] > 0
