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
@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("flags", [0, 31])
def test_c1_cpp_driver_full_trace_parity(flags, fused, tmp_path):
    """fused: the control steps between two observations go through cslam_ekf_control_steps (one
    single-CTA launch per batch) instead of one predict + one observeHeading call per step."""
    from conan_slam_b200 import build
    exe = build.build_host()
    trace = str(tmp_path / "trace.bin")
    out = subprocess.run([exe, "--flags", str(flags), "--trace", trace, "--print-every", "0"] +
                         (["--fused"] if fused else []), capture_output=True, text=True, timeout=600)
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


def _lcg_draws(seed, count):
    """The deterministic stand-in draw tape of conan_slam_b200/host/pf_main.cpp (lcg_normalish)."""
    out = np.empty(count)
    s = seed
    mask = (1 << 64) - 1
    for i in range(count):
        acc = 0.0
        for _ in range(4):
            s = (s * 6364136223846793005 + 1442695040888963407) & mask
            acc += float((s >> 11) & 0xFFFFFFFFFFFFF) / 4503599627370496.0
        out[i] = (acc - 2.0) * 1.7320508075688772
    return out, s


@pytest.mark.gpu
def test_pf_cpp_population_adaptor_matches_oracle(tmp_path):
    """PF half of test/main.cpp (:204-335) through the C++ adaptor PfGpuT (host/pf_main.cpp) against
    the CPU oracle driven through the same call sequence with the same draw tape."""
    from conan_slam_b200 import build
    build.build_host()
    exe = os.path.join(ROOT, "conan_slam_b200", "lib", "pf_main")
    P, cycles, flags = 512, 3, oracle_py.FLAG_INTENDED
    outf = str(tmp_path / "pf.bin")
    run = subprocess.run([exe, "--particles", str(P), "--cycles", str(cycles), "--flags", str(flags), "--out", outf],
                         capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "skipped=0" in run.stdout
    raw = np.fromfile(outf, dtype=np.float64)
    w_g, x_g, keep_g, neff_g = raw[:P], raw[P:4 * P].reshape(P, 3), raw[4 * P:5 * P].astype(np.int64), raw[5 * P]

    o = oracle_py.OraclePF(P, flags)
    Q = np.diag([2 * 0.3 ** 2, 2 * (np.pi / 180.0) ** 2])
    R = np.diag([2 * 0.1 ** 2, 2 * (np.pi / 180.0) ** 2])
    seed = 42

    def controls(c0):
        for c in range(6):
            o.predict(83.33, 0.02 * np.sin(0.3 * (c0 + c)), Q, 73.0, 0.01)
            o.observeHeading(0.001 * (c0 + c), True)

    controls(0)
    xi, seed = _lcg_draws(seed, 3 * P)
    o.samplePose(xi.reshape(P, 3))
    zr, zb = np.array([400.0, 900.0, 650.0]), np.array([0.3, -0.7, 0.05])
    o.addOneNewFeature(np.stack([zr, zb]), R)
    ids = np.array([1, 3], dtype=np.int32)
    keep_o = neff_o = None
    for c in range(cycles):
        controls(6 * (c + 1))
        ZF = np.array([[zr[0] - 5.0 * (c + 1), zr[2] - 5.0 * (c + 1)],
                       [zb[0] + 0.002 * (c + 1), zb[2] - 0.001 * (c + 1)]])
        xi, seed = _lcg_draws(seed, 3 * P)
        o.sampleProposal(ZF, ids, R, xi.reshape(P, 3))
        o.featureUpdate(ZF, ids, R)
        u, seed = _lcg_draws(seed, P)
        keep_o, neff_o, did = o.resampleParticles(1e300, u * 0.3, True)
        assert did
    assert np.array_equal(keep_g, keep_o)
    assert abs(neff_g - neff_o) < 1e-9 * neff_o
    assert rel_err(w_g, o.weights) < 1e-9
    assert rel_err(x_g, o.poses) < 1e-9
