"""GPU parity tests (through the C ABI) of the particle-filter hot path against the CPU oracle.

Bar: poses / features / weights within 1e-9 relative in FP64 (weights: products of Gaussian
densities through exp/pow, 1e-9 as well); resampled indices exactly equal given identical draws.
"""
import numpy as np
import pytest

import helpers
import oracle_py
from helpers import QE, rel_err

pytestmark = pytest.mark.gpu

R2 = 2 * helpers.R_BASE  # test/main.cpp:245
TOL = 1e-9


def _scenario(npart, nfeat, flags, seed, m_obs=3, good_cov=True):
    """Drives oracle and GPU through the reference's PF call sequence (test/main.cpp:204-335)
    with shared draws; returns both filters after sampleProposal + featureUpdate."""
    import conan_slam_b200 as cs
    rng = np.random.default_rng(seed)
    g = cs.PF(num_particles=npart, capacity_landmarks=nfeat + 4, flags=flags)
    o = oracle_py.OraclePF(npart, flags)
    lm = rng.uniform(-900, 900, size=(2, nfeat))
    for f in (g, o):
        for k in range(6):
            f.predict(83.33, 0.03, QE, 73.0, 0.01)
            f.observeHeading(0.0005 * (k + 1), True)
    X = o.poses
    Z0 = np.zeros((2, nfeat))
    for j in range(nfeat):
        Z0[:, j] = [np.hypot(lm[0, j] - X[0, 0], lm[1, j] - X[0, 1]),
                    np.arctan2(lm[1, j] - X[0, 1], lm[0, j] - X[0, 0]) - X[0, 2]]
    xi0 = rng.normal(size=(npart, 3))
    for f in (g, o):
        f.samplePose(xi0)                   # test/main.cpp:319-325
        f.addOneNewFeature(Z0, R2)          # test/main.cpp:326
        for k in range(6):
            f.predict(83.33, -0.02, QE, 73.0, 0.01)
            f.observeHeading(0.004 + 0.0005 * k, True)
    if good_cov:
        base = np.array([[4e-4, 1e-5, 1e-7], [1e-5, 5e-4, -2e-7], [1e-7, -2e-7, 3e-8]])
        covs = np.tile(base.reshape(-1), (npart, 1)) * (1.0 + 0.1 * rng.uniform(size=(npart, 1)))
        for f in (g, o):
            f.set_poses(o.poses, covs)
    poses = o.poses
    ids = (rng.choice(nfeat, size=m_obs, replace=False) + 1).astype(np.int32)
    Z = np.zeros((2, m_obs))
    for k, j in enumerate(ids):
        Z[:, k] = [np.hypot(lm[0, j - 1] - poses[0, 0], lm[1, j - 1] - poses[0, 1]) + 0.004,
                   np.arctan2(lm[1, j - 1] - poses[0, 1], lm[0, j - 1] - poses[0, 0]) - poses[0, 2] + 2e-6]
    xi = rng.normal(size=(npart, 3))
    for f in (g, o):
        f.sampleProposal(Z, ids, R2, xi)
        f.featureUpdate(Z, ids, R2)
    return g, o, rng


def _assert_particles(g, o, which):
    assert rel_err(g.poses, o.poses) < TOL
    assert rel_err(g.pose_covs, o.pose_covs) < TOL
    wg, wo = g.weights, o.weights
    assert np.all(np.isfinite(wo)) and np.all(wo > 0)
    assert np.max(np.abs(wg - wo) / wo) < TOL
    for p in which:
        XFg, PFg = g.features(p)
        XFo, PFo = o.features(p)
        assert rel_err(XFg, XFo) < TOL and rel_err(PFg, PFo) < TOL


@pytest.mark.parametrize("flags", [0, oracle_py.FLAG_INTENDED])
@pytest.mark.parametrize("npart,nfeat", [(1, 3), (33, 5), (1000, 12)])
def test_pf_step_parity(npart, nfeat, flags):
    g, o, rng = _scenario(npart, nfeat, flags, seed=npart + nfeat)
    assert g.num_features == nfeat == o.num_features
    _assert_particles(g, o, sorted(set([0, npart // 2, npart - 1])))
    assert g.sync() == 0
    Xe_g, ig = g.extractStatesFromParticles()
    Xe_o, io = o.extractStatesFromParticles()
    assert ig == io and rel_err(Xe_g, Xe_o) < TOL


@pytest.mark.parametrize("npart", [1, 31, 32, 33, 1000, 4097, 70000])
def test_resample_indices_exact_intended(npart):
    """INTENDED resampler: canonical radix-32 scan + first-true search; indices, neff and the
    gathered particle set must match the oracle exactly / to 1e-12."""
    import conan_slam_b200 as cs
    flags = oracle_py.FLAG_INTENDED
    rng = np.random.default_rng(npart)
    nfeat = 3 if npart > 5000 else 6
    g, o, _ = _scenario(npart, nfeat, flags, seed=1000 + npart, m_obs=2)
    # spread the weights a lot so neff drops and many particles are duplicated / dropped
    w = rng.uniform(0.0, 1.0, size=npart) ** 4 + 1e-12
    g.weights = w
    o.weights = w
    u = rng.normal(size=npart) * 0.3     # Q12: the reference draws NORMAL deviates for the comb
    kg, neff_g, did_g = g.resampleParticles(npart + 1, u, True)
    ko, neff_o, did_o = o.resampleParticles(npart + 1, u, True)
    assert did_g and did_o
    assert np.array_equal(kg, ko)
    assert neff_g == pytest.approx(neff_o, rel=1e-12)
    assert np.all(g.weights == 1.0 / npart)
    assert rel_err(g.poses, o.poses) < 1e-12        # gather-copy moved exactly the selected particles
    for p in sorted(set([0, npart // 3, npart - 1])):
        XFg, PFg = g.features(p)
        XFo, PFo = o.features(p)
        assert rel_err(XFg, XFo) < 1e-12 and rel_err(PFg, PFo) < 1e-12
    # resampling off / neff above threshold: weights normalised, particles untouched
    g.weights = w
    o.weights = w
    poses_before = g.poses
    kg2, neff2, did2 = g.resampleParticles(0.5, u, True)
    assert not did2 and np.array_equal(g.poses, poses_before)
    ws = o.weights
    o.resampleParticles(0.5, u, True)
    assert np.max(np.abs(g.weights - o.weights) / o.weights) < 1e-12


@pytest.mark.parametrize("case", ["one_mass", "uniform"])
def test_resample_degenerate_weights(case):
    import conan_slam_b200 as cs
    npart = 257
    g = cs.PF(num_particles=npart, capacity_landmarks=2, flags=oracle_py.FLAG_INTENDED)
    u = np.zeros(npart)
    if case == "one_mass":
        w = np.zeros(npart)
        w[101] = 5.0
        g.weights = w
        keep, neff, did = g.resampleParticles(npart, u, True)
        assert np.all(keep == 101) and neff == pytest.approx(1.0) and did
    else:
        g.weights = np.full(npart, 0.25)
        keep, neff, did = g.resampleParticles(npart * 0.75, u, True)
        assert np.array_equal(keep, np.arange(npart)) and neff == pytest.approx(npart) and not did
        ko, _, _ = oracle_py.stratified_resample(np.full(npart, 0.25), u, oracle_py.FLAG_INTENDED)
        assert np.array_equal(keep, ko)


@pytest.mark.parametrize("npart", [1, 20, 300])
def test_resample_literal_q10(npart):
    """REF_LITERAL (PF.cpp:566-574): the first slot whose comb value is below its own
    cumulative weight takes every slot; sums are plain left-to-right."""
    import conan_slam_b200 as cs
    rng = np.random.default_rng(npart)
    g = cs.PF(num_particles=npart, capacity_landmarks=2, flags=0)
    w = rng.uniform(0.05, 1.0, size=npart)
    u = rng.normal(size=npart) * 0.2
    g.weights = w
    kg, neff_g, _ = g.resampleParticles(0.0, u, False)
    o = oracle_py.OraclePF(npart, 0)
    o.weights = w
    ko, neff_o, _ = o.resampleParticles(0.0, u, False)
    assert np.array_equal(kg, ko) and len(set(kg.tolist())) == 1
    assert neff_g == pytest.approx(neff_o, rel=1e-12)


def test_pf_argument_errors():
    import conan_slam_b200 as cs
    g = cs.PF(num_particles=16, capacity_landmarks=2)
    xi = np.zeros((16, 3))
    g.addOneNewFeature(np.array([[10.0, 20.0], [0.1, 0.2]]), R2)
    with pytest.raises(cs.CslamError) as e:
        g.addOneNewFeature(np.array([[10.0], [0.1]]), R2)
    assert e.value.code == 2
    with pytest.raises(cs.CslamError) as e:
        g.featureUpdate(np.array([[10.0], [0.1]]), np.array([3], dtype=np.int32), R2)
    assert e.value.code == 1
    with pytest.raises(cs.CslamError) as e:
        g.sampleProposal(np.array([[10.0, 11.0], [0.1, 0.1]]), np.array([1, 1], dtype=np.int32), R2, xi)
    assert e.value.code == 5


@pytest.mark.parametrize("use_heading", [True, False])
def test_control_steps_equal_stepwise_calls(use_heading):
    """cslam_pf_control_steps (k x predict + observeHeading in registers, one launch) is bit-identical to
    the 2k per-step launches and matches the oracle."""
    import conan_slam_b200 as cs
    npart = 1000
    rng = np.random.default_rng(3)
    a = cs.PF(num_particles=npart, capacity_landmarks=2, flags=oracle_py.FLAG_INTENDED)
    b = cs.PF(num_particles=npart, capacity_landmarks=2, flags=oracle_py.FLAG_INTENDED)
    o = oracle_py.OraclePF(npart, oracle_py.FLAG_INTENDED)
    X0 = rng.normal(size=(npart, 3)) * np.array([5.0, 5.0, 0.05])
    for f in (a, b, o):
        f.set_poses(X0, None)
    k = 19  # > 16: two launches
    v = 83.33 + rng.normal(size=k)
    swa = 0.05 * rng.normal(size=k)
    phi = np.cumsum(v * 0.01 * np.sin(swa) / 73.0) + 1e-4 * rng.normal(size=k)
    a.controlSteps(v, swa, phi, use_heading, QE, 73.0, 0.01)
    for i in range(k):
        for f in (b, o):
            f.predict(v[i], swa[i], QE, 73.0, 0.01)
            f.observeHeading(phi[i], use_heading)
    assert np.array_equal(a.poses, b.poses) and np.array_equal(a.pose_covs, b.pose_covs)
    assert rel_err(a.poses, o.poses) < TOL


def test_checkpoint_roundtrip(tmp_path):
    """cslam_pf_save / _load: a restored particle set continues bit-identically."""
    import conan_slam_b200 as cs
    g, o, rng = _scenario(600, 5, oracle_py.FLAG_INTENDED, seed=21)
    path = tmp_path / "pf.ckpt"
    g.save(path)
    g2 = cs.PF(num_particles=600, capacity_landmarks=9, flags=oracle_py.FLAG_INTENDED)
    g2.load(path)
    assert g2.num_features == g.num_features
    assert np.array_equal(g2.weights, g.weights) and np.array_equal(g2.poses, g.poses)
    assert np.array_equal(g2.pose_covs, g.pose_covs)
    for p in (0, 17, 599):
        a, b = g.features(p), g2.features(p)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    u = rng.normal(size=600) * 0.3
    ka, na, da = g.resampleParticles(600, u, True)
    kb, nb, db = g2.resampleParticles(600, u, True)
    assert np.array_equal(ka, kb) and na == nb and da == db
    assert np.array_equal(g2.poses, g.poses)
    other = cs.PF(num_particles=512, capacity_landmarks=9)
    with pytest.raises(Exception):
        other.load(path)  # particle count mismatch


def test_resample_five_level_scan_hierarchy():
    """More than 32^4 = 1,048,576 particles: the canonical scan needs a fifth level (round-1 advice: the host loop
    read one past the end of its level table there).  Indices and neff against the oracle's stratifiedResample."""
    import conan_slam_b200 as cs
    npart = (1 << 20) + 96
    rng = np.random.default_rng(77)
    g = cs.PF(num_particles=npart, capacity_landmarks=1, flags=oracle_py.FLAG_INTENDED)
    w = rng.uniform(0.0, 1.0, size=npart) ** 3 + 1e-12
    u = rng.normal(size=npart) * 0.3
    o = oracle_py.OraclePF(npart, oracle_py.FLAG_INTENDED)
    g.weights = w
    o.weights = w
    kg, neff_g, did = g.resampleParticles(npart + 1, u, True)
    ko, neff_o, did_o = o.resampleParticles(npart + 1, u, True)
    assert did and did_o and np.array_equal(kg, ko)
    assert neff_g == pytest.approx(neff_o, rel=1e-12)
    g.close()
