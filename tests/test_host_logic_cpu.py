"""CPU tests of host-side logic that needs no GPU: the ragged trace container, the bench's
pure helpers and the reference arm's JSON contract (run on a tiny workload)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ragged_trace_indexing():
    from zfista_b200.proximal_gradient import RaggedTrace

    lens = np.array([3, 0, 5])
    off = np.zeros(4, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    errs = RaggedTrace(np.arange(8.0), off)
    assert len(errs) == 3
    np.testing.assert_array_equal(errs[0], [0, 1, 2])
    assert errs[1].shape == (0,)
    np.testing.assert_array_equal(errs[2, :2], [3, 4])
    np.testing.assert_array_equal(errs[2][-1:], [7])
    # F / x traces have one more entry per start (entry 0 = the start itself)
    foff = off + np.arange(4)
    funs = RaggedTrace(np.arange(11 * 2.0).reshape(11, 2), foff)
    assert funs[0].shape == (4, 2) and funs[1].shape == (1, 2) and funs[2].shape == (6, 2)
    np.testing.assert_array_equal(funs[1][0], [8, 9])
    np.testing.assert_array_equal(funs[2, :2], [[10, 11], [12, 13]])
    assert [a.shape[0] for a in funs] == [4, 1, 6]
    assert funs.nbytes == 11 * 2 * 8


def test_bench_helpers():
    sys.path.insert(0, ROOT)
    import bench

    h = bench.nit_histogram(np.array([10, 60, 150, 250, 250, 590]))
    assert sum(h["counts"]) == 6 and h["max"] == 590
    assert abs(h["mean_over_max_utilisation"] - np.mean([10, 60, 150, 250, 250, 590]) / 590) < 1e-12
    spec = bench.workload_spec("fds")
    cfg = bench.headline_config(spec)
    assert cfg["options"]["max_iter_internal"] == 100 and "options_note" in cfg
    assert "options_note" not in bench.headline_config(bench.workload_spec("jos1"))
    r = bench.fp64_roofline(3.0e-3, 260000, None)
    assert r["bound"] == "fp64" and r["unit"] == "TFLOP/s" and 0 < r["frac"] < 1
    assert abs(r["peak"] - 148 * 64 * 2 * 1.965e9 / 1e12) < 1e-9


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    """bench.py --impl reference on the cheap JOS1 workload (the FDS one is minutes of
    trust-constr): one JSON line on stdout, same config object as the GPU arm, steps completed,
    every solve counted."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--workload", "jos1", "--steps", "2", "--warmup", "1", "--ref-sample", "8"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1
    assert d["solves_timed"] == 16 and d["converged_fraction"] == 1.0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("JOS1") and d["gpu_launches"] == 0
