"""Config 1 (BASELINE.json configs[0]): the reference's test/main.cpp waypoint loop replayed by the
C++ host driver (conan_slam_b200/host/sim_main.cpp -> slam_gpu.hpp -> C ABI -> CUDA) against the
CPU oracle replaying the same tape.  X within 1e-9 relative after EVERY control step."""
import os
import subprocess

import numpy as np
import pytest

import oracle_py
from helpers import rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "c1_trace_literal.npz")


def _oracle_trace(flags, last=None):
    tape = oracle_py.sim_tape(noise_seed=0)
    o = oracle_py.OracleEKF(flags)
    o.reset(np.zeros(3), np.zeros((3, 3)))
    rows = []
    oracle_py.run_tape(o, tape, last=last, dense_heading=False,
                       on_step=lambda s, f: rows.append(np.concatenate([f.X[:3], [f.n]])))
    return tape, np.asarray(rows), o.X


def test_c1_oracle_matches_golden_fixture():
    """The committed golden trace (tests/golden/make_c1_golden.py) pins the oracle's config-1 run."""
    g = np.load(GOLD)
    tape, rows, Xf = _oracle_trace(0)
    assert tape["steps"] == int(g["steps"])
    assert np.array_equal(rows[:: int(g["stride"])][:, 3], g["trace"][:, 3])
    assert rel_err(rows[:: int(g["stride"])][:, :3], g["trace"][:, :3]) < 1e-12
    assert rel_err(Xf, g["X_final"]) < 1e-12
    # the estimate tracks the (noise-free) truth and ends at the last waypoint
    assert Xf.shape[0] == 53
    assert abs(Xf[0] - (-4976.478)) < 1.5 and abs(Xf[1] - 1464.968) < 1.5


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [0, 31])
def test_c1_cpp_driver_full_trace_parity(flags, tmp_path):
    from conan_slam_b200 import build
    exe = build.build_host()
    trace = str(tmp_path / "trace.bin")
    out = subprocess.run([exe, "--flags", str(flags), "--trace", trace, "--print-every", "0"], capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    assert "skipped updates=0" in out.stdout
    tape, rows, Xf = _oracle_trace(flags)
    raw = np.fromfile(trace, dtype=np.float64)
    steps = tape["steps"]
    got = raw[: 4 * steps].reshape(steps, 4)
    Xg = raw[4 * steps:]
    assert np.array_equal(got[:, 3], rows[:, 3])            # map growth at identical steps
    assert rel_err(got[:, :2], rows[:, :2]) < 1e-9           # position, every control step
    assert np.max(np.abs(got[:, 2] - rows[:, 2])) < 1e-9     # heading (radians, |phi| <= pi)
    assert Xg.shape == Xf.shape and rel_err(Xg, Xf) < 1e-9   # final pose + all 25 landmarks
