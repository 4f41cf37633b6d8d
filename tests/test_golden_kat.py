"""Committed known-answer vectors (tests/golden/kat_r01.npz, generator tests/golden/make_kat_golden.py).
CPU tier: the oracle reproduces them (pins the oracle against regressions).  GPU tier: the CUDA path,
driven through the C ABI with the fixture's stored inputs, reproduces them — FP64 within 1e-9 relative,
association and resampling indices exactly."""
import os
import sys

import numpy as np
import pytest

import helpers
import oracle_py
from helpers import rel_err

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "kat_r01.npz"))
MODES = [("lit", 0), ("int", oracle_py.FLAG_INTENDED)]


def _g(prefix, name):
    return GOLD[f"{prefix}_{name}"]


def _run_ekf(f, pre):
    """The fixed call sequence of make_kat_golden.ekf_case on filter f, inputs taken from the fixture."""
    out = {}
    X0, P0 = _g(pre, "X0"), _g(pre, "P0")
    f.reset(X0, P0)
    f.predict(83.33, 0.02, helpers.QE, 73.0, 0.01)
    f.observeHeading(float(X0[2]) + 1e-4, True)
    out["X_ph"], out["P_ph"] = f.X.copy(), f.P.copy()
    g = f.gate(_g(pre, "Zall"), helpers.RE, 50.0, 1000.0)
    out["jbest"], out["is_new"], out["nbest"], out["outer"] = g[0], g[1], g[2], g[3]
    f.update(_g(pre, "Z"), helpers.RE, _g(pre, "ids"), False)
    out["X_seq"], out["P_seq"] = f.X.copy(), f.P.copy()
    f.update(_g(pre, "Z2"), helpers.RE, _g(pre, "ids2"), True)
    out["X_batch"], out["P_batch"] = f.X.copy(), f.P.copy()
    f.augment(_g(pre, "Zn"), helpers.RE)
    out["X_aug"], out["P_aug"] = f.X.copy(), f.P.copy()
    return out


def _check_ekf(out, pre, tol):
    assert np.array_equal(out["jbest"], _g(pre, "jbest")) and np.array_equal(out["is_new"], _g(pre, "is_new"))
    hit = _g(pre, "jbest") != 0
    assert rel_err(np.asarray(out["nbest"])[hit], _g(pre, "nbest")[hit]) < tol
    assert np.allclose(np.asarray(out["outer"])[~hit], _g(pre, "outer")[~hit], rtol=tol, atol=0)
    for k in ("X_ph", "X_seq", "X_batch", "X_aug"):
        assert rel_err(out[k], _g(pre, k)) < tol, k
    for k in ("P_ph", "P_seq", "P_batch", "P_aug"):
        iu = np.triu_indices(_g(pre, k).shape[0])
        assert rel_err(np.asarray(out[k])[iu], _g(pre, k)[iu]) < tol, k


def _run_pf(f, pre):
    R2 = 2 * helpers.R_BASE
    for k in range(6):
        f.predict(83.33, 0.03, helpers.QE, 73.0, 0.01)
        f.observeHeading(0.0005 * (k + 1), True)
    f.samplePose(_g(pre, "xi0"))
    f.addOneNewFeature(_g(pre, "Z0"), R2)
    for k in range(6):
        f.predict(83.33, -0.02, helpers.QE, 73.0, 0.01)
        f.observeHeading(0.004 + 0.0005 * k, True)
    f.set_poses(f.poses, _g(pre, "covs"))
    f.sampleProposal(_g(pre, "Z"), _g(pre, "ids"), R2, _g(pre, "xi"))
    f.featureUpdate(_g(pre, "Z"), _g(pre, "ids"), R2)
    w_before = np.asarray(f.weights).copy()
    f.weights = _g(pre, "w_skew")
    keep, neff, did = f.resampleParticles(float(len(w_before)), _g(pre, "u"), True)
    return {"w_before": w_before, "keep": np.asarray(keep), "neff": neff, "did": did, "poses": np.asarray(f.poses),
            "weights": np.asarray(f.weights), "xf0": f.features(0)[0], "pf0": f.features(0)[1]}


def _check_pf(out, pre, tol):
    assert np.array_equal(out["keep"], _g(pre, "keep")) and int(out["did"]) == int(_g(pre, "did"))
    assert abs(out["neff"] - float(_g(pre, "neff"))) < tol * float(_g(pre, "neff"))
    for k in ("w_before", "poses", "weights", "xf0", "pf0"):
        assert rel_err(out[k], _g(pre, k)) < tol, k


@pytest.mark.parametrize("name,flags", MODES)
def test_oracle_reproduces_golden_kat(name, flags):
    _check_ekf(_run_ekf(oracle_py.OracleEKF(flags), f"ekf_{name}"), f"ekf_{name}", 1e-12)
    _check_pf(_run_pf(oracle_py.OraclePF(16, flags), f"pf_{name}"), f"pf_{name}", 1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("name,flags", MODES)
def test_gpu_reproduces_golden_kat(name, flags):
    import conan_slam_b200 as cs
    _check_ekf(_run_ekf(cs.EKF(capacity_landmarks=10, flags=flags), f"ekf_{name}"), f"ekf_{name}", 1e-9)
    _check_pf(_run_pf(cs.PF(num_particles=16, capacity_landmarks=4, flags=flags), f"pf_{name}"), f"pf_{name}", 1e-9)
