"""GPU parity tests (through the C ABI) of the EKF-SLAM hot path against the CPU oracle.

Bar (BASELINE.json north_star): state / landmark / covariance estimates within 1e-9 relative
in FP64; association indices exactly equal.  Inputs are seeded; sizes are what the dense oracle
finishes in seconds; the benchmarked sizes (2,000 landmarks with the full covariance, 20,000 landmarks on the
marginal of the observed landmarks) are covered by tests/test_ekf_lazy.py.
"""
import numpy as np
import pytest

import helpers
import oracle_py
from helpers import QE, RE, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-9  # north_star tolerance, relative to the largest entry of the compared array


def _pair(N, seed, flags, capacity=None):
    import conan_slam_b200 as cs
    X, P, lm = helpers.synthetic_map(N, seed)
    g = cs.EKF(capacity_landmarks=capacity or (N + 8), flags=flags)
    o = oracle_py.OracleEKF(flags)
    g.reset(X, P)
    o.reset(X, P)
    return g, o, lm


def _assert_state(g, o, tol=TOL):
    assert g.n == o.n
    assert rel_err(g.X, o.X) < tol
    Pg, Po = g.P, o.P
    assert np.array_equal(Pg, Pg.T)  # accessor mirrors the authoritative upper triangle
    iu = np.triu_indices(o.n)
    assert rel_err(Pg[iu], Po[iu]) < tol


@pytest.mark.parametrize("flags", [0, oracle_py.FLAG_INTENDED])
@pytest.mark.parametrize("N", [0, 1, 2, 25, 150])
def test_predict_heading_parity(N, flags):
    g, o, lm = _pair(N, 10 + N, flags)
    for k in range(3):
        for f in (g, o):
            f.predict(83.33 + k, 0.03 * (k - 1), QE, 73.0, 0.01)
        _assert_state(g, o)
        for f in (g, o):
            f.observeHeading(o.X[2] + 1e-4 * (k + 1), True)
        _assert_state(g, o)
    assert g.sync() == 0
    # useHeading = false is a no-op (EKF.cpp:332-335)
    before = g.X
    g.observeHeading(1.0, False)
    assert np.array_equal(g.X, before)


@pytest.mark.parametrize("flags", [0, oracle_py.FLAG_INTENDED])
@pytest.mark.parametrize("N,m", [(1, 1), (2, 2), (25, 5), (150, 7), (150, 20), (1100, 4), (1100, 9)])
def test_single_update_parity(N, m, flags):
    g, o, lm = _pair(N, 20 + N, flags)
    rng = np.random.default_rng(N)
    ids = (rng.choice(N, size=m, replace=False) + 1).astype(np.int32)
    Z = helpers.observe(o.X, lm, ids, rng)
    for f in (g, o):
        f.update(Z, RE, ids, False)
    _assert_state(g, o)
    assert g.sync() == 0


def test_sequential_update_large_map():
    """A larger map (n = 8 203, 128-wide tiles, several tile rows): fused scan = gate + gain kernels
    seeing P through the pending rank-2 terms + one multi pass, the path the 20 000-landmark bench takes."""
    g, o, lm = _pair(4100, 4120, oracle_py.FLAG_INTENDED)
    rng = np.random.default_rng(4100)
    ids = (rng.choice(4100, size=3, replace=False) + 1).astype(np.int32)
    Z = helpers.observe(o.X, lm, ids, rng)
    jg, _ = g.scan(Z, RE, 50.0, 1000.0)
    o.update(Z, RE, ids, False)
    assert np.array_equal(jg, ids)
    assert rel_err(g.X, o.X) < TOL
    # covariance: rows 0..2, the observed landmarks' rows and a random sample of the rest
    rows = np.unique(np.concatenate([[0, 1, 2], 3 + 2 * (ids - 1), 4 + 2 * (ids - 1), rng.choice(o.n, size=40)]))
    Pg = g.P
    for i in rows:
        assert rel_err(Pg[i, i:], o.P[i, i:]) < TOL
    assert g.sync() == 0


@pytest.mark.parametrize("flags", [0, oracle_py.FLAG_INTENDED])
@pytest.mark.parametrize("N,m", [(1, 1), (7, 7), (25, 5), (150, 16), (150, 32), (1100, 32), (1100, 5), (700, 17)])
def test_batch_update_parity(N, m, flags):
    """EKF.cpp:93-129 joint update; N=1100 takes the FP64 tensor-core (DMMA) kernel."""
    g, o, lm = _pair(N, 30 + N, flags)
    rng = np.random.default_rng(N + m)
    ids = (rng.choice(N, size=m, replace=False) + 1).astype(np.int32)
    Z = helpers.observe(o.X, lm, ids, rng)
    for f in (g, o):
        f.update(Z, RE, ids, True)
    _assert_state(g, o)
    assert g.sync() == 0


def test_empty_update_is_noop():
    """test/main.cpp:188 calls update with an empty ZF whenever every visible landmark is new."""
    g, o, lm = _pair(5, 3, 0)
    before_x, before_p = g.X, g.P
    g.update(np.zeros((2, 0)), RE, np.zeros(0, dtype=np.int32), True)
    g.update(np.zeros((2, 0)), RE, np.zeros(0, dtype=np.int32), False)
    assert np.array_equal(g.X, before_x) and np.array_equal(g.P, before_p)


@pytest.mark.parametrize("flags", [0, oracle_py.FLAG_INTENDED])
def test_augment_parity_from_empty_map(flags):
    """EKF.cpp:9-91: growth from n = 3 (no landmarks) inside the pre-allocated covariance."""
    import conan_slam_b200 as cs
    g = cs.EKF(capacity_landmarks=40, flags=flags)
    o = oracle_py.OracleEKF(flags)
    X0 = np.array([3.0, -2.0, 0.4])
    P0 = np.array([[0.5, 0.02, 0.001], [0.02, 0.4, -0.002], [0.001, -0.002, 3e-4]])
    g.reset(X0, P0)
    o.reset(X0, P0)
    rng = np.random.default_rng(0)
    for round_ in range(4):
        m = [1, 3, 7, 9][round_]
        Z = np.stack([rng.uniform(50, 1900, size=m), rng.uniform(-1.5, 1.5, size=m)])
        for f in (g, o):
            f.augment(Z, RE)
            f.predict(83.33, 0.01, QE, 73.0, 0.01)
        _assert_state(g, o)
    assert g.num_landmarks == 20


def test_augment_capacity_error():
    import conan_slam_b200 as cs
    g = cs.EKF(capacity_landmarks=2)
    g.reset(np.zeros(3), np.eye(3))
    g.augment(np.array([[10.0, 20.0], [0.1, 0.2]]), RE)
    with pytest.raises(cs.CslamError) as e:
        g.augment(np.array([[30.0], [0.3]]), RE)
    assert e.value.code == 2
    with pytest.raises(cs.CslamError) as e:
        g.update(np.array([[30.0], [0.3]]), RE, np.array([3], dtype=np.int32), False)  # idf out of range
    assert e.value.code == 1


@pytest.mark.parametrize("N,m", [(1, 1), (3, 4), (40, 6), (600, 64), (2000, 4)])
def test_gating_indices_exact(N, m):
    """EKF.cpp:235-326: association indices must equal the oracle's exactly; the test also
    checks the decisions are well separated (no accidental near-ties)."""
    g, o, lm = _pair(N, 50 + N, 0)
    rng = np.random.default_rng(7 * N + m)
    k_assoc = min(N, max(1, m // 2))
    ids = rng.choice(N, size=k_assoc, replace=False) + 1
    Z = helpers.observe(o.X, lm, ids, rng)
    # the rest: spurious observations (mostly far from every landmark)
    extra = m - k_assoc
    if extra > 0:
        Zx = np.stack([rng.uniform(30.0, 3000.0, size=extra), rng.uniform(-np.pi, np.pi, size=extra)])
        Z = np.concatenate([Z, Zx], axis=1)
    jg, newg, nbg, outg = g.gate(Z, RE, 50.0, 1000.0)
    jo, newo, nbo, outo, idf_o, _ = o.gate(Z, RE, 50.0, 1000.0, dense=(N <= 40))
    assert np.array_equal(jg, jo)
    assert np.array_equal(newg, newo)
    assert np.array_equal(jg[:k_assoc], ids)
    hit = jo != 0
    assert rel_err(nbg[hit], nbo[hit]) < 1e-12 if hit.any() else True
    miss = ~hit
    if miss.any():
        assert np.allclose(outg[miss], outo[miss], rtol=1e-12, atol=0)
    # dataAssociate wrapper: ZF / idf in observation order; ZN empty in REF_LITERAL (Q5)
    a = g.dataAssociate(Z, RE, 50.0, 1000.0)
    assert np.array_equal(a.idf, idf_o)
    assert np.array_equal(a.ZF, Z[:, hit])
    assert a.ZN.size == 0


@pytest.mark.parametrize("flags", [0, oracle_py.FLAG_INTENDED])
@pytest.mark.parametrize("N,m", [(0, 2), (3, 4), (40, 6), (600, 9), (600, 21), (1100, 5)])
def test_fused_scan_equals_gate_then_update(N, m, flags):
    """cslam_ekf_scan (association indices stay on the device) == dataAssociate + singleUpdate
    (test/main.cpp:193-195) of the oracle: same indices, same state; spurious observations that
    pass no gate must leave the filter untouched."""
    g, o, lm = _pair(N, 70 + N, flags)
    rng = np.random.default_rng(11 * N + m)
    k_assoc = min(N, max(1, m // 2)) if N else 0
    ids = rng.choice(N, size=k_assoc, replace=False) + 1 if N else np.zeros(0, dtype=np.int64)
    Z = helpers.observe(o.X, lm, ids, rng) if k_assoc else np.zeros((2, 0))
    extra = m - k_assoc
    Zx = np.stack([rng.uniform(3000.0, 9000.0, size=extra), rng.uniform(-np.pi, np.pi, size=extra)])
    Z = np.concatenate([Z, Zx], axis=1)
    Z = Z[:, rng.permutation(m)]  # associated and spurious observations interleaved
    total = 0
    for rep in range(2):  # second scan: asynchronous form, indices fetched by a separate gate
        jo, newo, _, _, idf_o, _ = o.gate(Z, RE, 50.0, 1000.0, dense=(N <= 40))
        o.update(Z[:, jo != 0], RE, idf_o, False)
        if rep == 0:
            jg, newg = g.scan(Z, RE, 50.0, 1000.0)
            assert np.array_equal(jg, jo) and np.array_equal(newg, newo)
        else:
            assert g.scan(Z, RE, 50.0, 1000.0, want_indices=False) is None
        _assert_state(g, o)
        total += int((jo != 0).sum())
        assert g.scan_associations() == total  # device-side count of applied updates
    assert g.sync() == 0


def test_gating_tie_lowest_index_wins_and_no_landmarks():
    import conan_slam_b200 as cs
    X = np.array([0.0, 0.0, 0.0, 100.0, 50.0, 100.0, 50.0, 100.0, 50.0])
    P = np.zeros((9, 9))
    P[0:3, 0:3] = np.diag([0.5, 0.5, 1e-4])
    blk = np.array([[2.0, 0.3], [0.3, 1.0]])
    cross = np.array([[0.1, 0.0], [0.0, 0.1], [0.001, -0.002]])
    for k in range(3):
        s = 3 + 2 * k
        P[s:s + 2, s:s + 2] = blk
        P[0:3, s:s + 2] = cross
        P[s:s + 2, 0:3] = cross.T
    g = cs.EKF(capacity_landmarks=4)
    g.reset(X, P)
    Z = np.array([[np.hypot(100, 50) + 0.05], [np.arctan2(50, 100) + 0.001]])
    assert g.gate(Z, RE, 50.0, 1000.0)[0][0] == 1          # three exact ties -> lowest j (Q4)
    g.reset(np.zeros(3), np.eye(3))                          # nf = 0: everything is "new"
    j, new, nb, out = g.gate(Z, RE, 50.0, 1000.0)
    assert j[0] == 0 and new[0] == 1 and np.isinf(out[0]) and np.isinf(nb[0])
    a = cs.EKF(capacity_landmarks=4, flags=cs.FLAG_INTENDED)
    a.reset(np.zeros(3), np.eye(3))
    assert a.dataAssociate(Z, RE, 50.0, 1000.0).ZN.shape == (2, 1)  # Q5 intended: new features returned


def test_non_spd_update_is_skipped_and_counted():
    """slam.h:252-255: a non-finite Cholesky inverse means zero gain — state unchanged."""
    import conan_slam_b200 as cs
    X = np.array([0.0, 0.0, 0.0, 50.0, 10.0])
    P = -np.eye(5)  # negative definite -> S = H P H^T + R not SPD
    g = cs.EKF(capacity_landmarks=2)
    o = oracle_py.OracleEKF(0)
    g.reset(X, P)
    o.reset(X, P)
    Z = np.array([[51.0], [0.2]])
    Rsmall = np.diag([1e-3, 1e-6])
    g.update(Z, Rsmall, np.array([1], dtype=np.int32), False)
    assert o.update(Z, Rsmall, np.array([1], dtype=np.int32), False) == 1
    assert g.sync() == 1
    assert np.array_equal(g.X, X)
    assert np.array_equal(g.P, P)
    g.update(Z, Rsmall, np.array([1], dtype=np.int32), True)
    assert g.sync() == 2 and np.array_equal(g.X, X)


def test_bearing_wrap_at_pi():
    """Innovation bearing is wrapped through pi2Pi (EKF.cpp:471): observe a landmark almost
    straight behind the vehicle so z_b - zhat_b crosses +-pi."""
    import conan_slam_b200 as cs
    X = np.array([0.0, 0.0, 0.0, -100.0, 0.5])
    P = np.diag([0.2, 0.2, 1e-4, 1.0, 1.0])
    g = cs.EKF(capacity_landmarks=2)
    o = oracle_py.OracleEKF(0)
    for f in (g, o):
        f.reset(X, P)
    zb = np.arctan2(-0.5, -100.0)  # just below -pi side, while zhat is just below +pi
    Z = np.array([[100.0], [zb]])
    for f in (g, o):
        f.update(Z, RE, np.array([1], dtype=np.int32), False)
    assert rel_err(g.X, o.X) < TOL
    assert abs(g.X[4] - 0.5) < 1.0  # wrapped innovation: a small correction, not a 2*pi swing


def test_main_loop_prefix_parity_literal():
    """Config 1 (test/main.cpp loop, known associations, batch update, heading known),
    first 1500 control steps with the LITERAL dense Joseph update in the oracle."""
    import conan_slam_b200 as cs
    tape = oracle_py.sim_tape(noise_seed=0)
    g = cs.EKF(capacity_landmarks=30, flags=0)
    o = oracle_py.OracleEKF(0)
    for f in (g, o):
        f.reset(np.zeros(3), np.zeros((3, 3)))
    tg = oracle_py.run_tape(g, tape, last=1500)
    to = oracle_py.run_tape(o, tape, last=1500, dense_heading=True)
    assert np.array_equal(tg, to)
    assert g.n == o.n and g.n > 3
    assert rel_err(g.X, o.X) < TOL
    iu = np.triu_indices(o.n)
    assert rel_err(g.P[iu], o.P[iu]) < TOL


def test_checkpoint_roundtrip_and_landmark_marginals(tmp_path):
    """SURVEY §8f "next" rows: checkpoint of X + upper triangle of P restores a bit-identical filter
    (same subsequent updates); landmark 2x2 marginals come back without reading P."""
    import conan_slam_b200 as cs
    g, o, lm = _pair(60, 321, oracle_py.FLAG_INTENDED)
    rng = np.random.default_rng(5)
    ids = (rng.choice(60, size=4, replace=False) + 1).astype(np.int32)
    Z = helpers.observe(o.X, lm, ids, rng)
    g.update(Z, RE, ids, False)
    Pg = g.P
    covs = g.landmark_covs()
    assert covs.shape == (60, 2, 2)
    for j in (0, 17, 59):
        f = 3 + 2 * j
        assert np.array_equal(covs[j], Pg[f:f + 2, f:f + 2])
    assert np.array_equal(g.landmark_covs(first=10, count=3), covs[9:12])
    path = tmp_path / "ekf.ckpt"
    g.save(path)
    # header + X + rows 0..2 (3 x n) + rows >= 3 from the diagonal on
    assert path.stat().st_size == 32 + 8 * g.n + 24 * g.n + 4 * (g.n - 3) * (g.n - 2)
    g2 = cs.EKF(capacity_landmarks=64, flags=oracle_py.FLAG_INTENDED)
    g2.load(path)
    assert g2.n == g.n and np.array_equal(g2.X, g.X) and np.array_equal(g2.P, Pg)
    for f in (g, g2):
        f.predict(83.33, 0.01, helpers.QE, 73.0, 0.01)
        f.observeHeading(float(o.X[2]) + 1e-4, True)
        f.update(Z[:, :2], RE, ids[:2], True)
    assert np.array_equal(g2.X, g.X) and np.array_equal(g2.P, g.P)
    small = cs.EKF(capacity_landmarks=8)
    with pytest.raises(Exception):
        small.load(path)  # exceeds capacity


@pytest.mark.parametrize("flags", [0, oracle_py.FLAG_INTENDED])
@pytest.mark.parametrize("N", [0, 30, 400, 700])
def test_control_steps_equal_stepwise_calls(N, flags):
    """cslam_ekf_control_steps (single-CTA launch for n <= 1024, per-step kernels above) is bit-identical
    to k x (predict, observeHeading) and matches the oracle."""
    import conan_slam_b200 as cs
    X, P, lm = helpers.synthetic_map(N, 40 + N)
    a = cs.EKF(capacity_landmarks=N + 2, flags=flags)
    b = cs.EKF(capacity_landmarks=N + 2, flags=flags)
    o = oracle_py.OracleEKF(flags)
    for f in (a, b, o):
        f.reset(X, P)
    k = 19  # > kMaxControlSteps: two launches
    rng = np.random.default_rng(N)
    v = 83.33 + rng.normal(size=k)
    swa = 0.05 * rng.normal(size=k)
    phi = X[2] + np.cumsum(v * 0.01 * np.sin(swa) / 73.0) + 1e-4 * rng.normal(size=k)
    trace = a.controlSteps(v, swa, phi, True, QE, 73.0, 0.01)
    want = []
    for i in range(k):
        for f in (b, o):
            f.predict(v[i], swa[i], QE, 73.0, 0.01)
            f.observeHeading(phi[i], True)
        want.append(b.X[:3].copy())
    assert np.array_equal(trace, np.asarray(want))
    assert np.array_equal(a.X, b.X) and np.array_equal(a.P, b.P)
    _assert_state(a, o)
    # heading unknown: predict only
    a.controlSteps(v[:3], swa[:3], phi[:3], False, QE, 73.0, 0.01, want_trace=False)
    for i in range(3):
        b.predict(v[i], swa[i], QE, 73.0, 0.01)
    assert np.array_equal(a.X, b.X) and np.array_equal(a.P, b.P)


@pytest.mark.parametrize("N", [62, 63, 126, 127, 510, 1022, 1023, 1086, 1087])
def test_tile_boundary_sizes(N):
    """State sizes straddling the tile edges of every covariance kernel (n = 3 + 2N around 128, 256,
    1024 = DMMA threshold, 2048 = switch from 64- to 128-wide tiles, 17 x 128): heading update (rank 1),
    fused scan (grouped rank-2 pass), joint update (FMA kernel below n = 1024, tensor cores above),
    augmentation into the last tile — full covariance against the oracle."""
    g, o, lm = _pair(N, 7000 + N, oracle_py.FLAG_INTENDED, capacity=N + 3)
    rng = np.random.default_rng(N)
    phi1 = float(o.X[2]) + 1e-4
    for f in (g, o):
        f.predict(83.33, 0.01, QE, 73.0, 0.01)
        f.observeHeading(phi1, True)
    _assert_state(g, o)
    ids = (rng.choice(N, size=5, replace=False) + 1).astype(np.int32)
    Z = helpers.observe(o.X, lm, ids, rng)
    jg, _ = g.scan(Z, RE, 50.0, 1000.0)
    assert np.array_equal(jg, ids)
    o.update(Z, RE, ids, False)
    _assert_state(g, o)
    ids2 = (rng.choice(N, size=3, replace=False) + 1).astype(np.int32)
    Z2 = helpers.observe(o.X, lm, ids2, rng)
    Zn = np.array([[650.0, 900.0, 410.0], [0.3, -0.8, 1.7]])
    phi2 = float(o.X[2]) - 5e-5
    for f in (g, o):
        f.update(Z2, RE, ids2, True)
        f.augment(Zn, RE)
        f.observeHeading(phi2, True)
        f.update(Zn[:, 2:], RE, np.array([N + 3], dtype=np.int32), False)
    _assert_state(g, o)
    assert g.sync() == 0


@pytest.mark.gpu
def test_observe_step_equals_update_then_augment():
    """cslam_ekf_observe_step: on a small map the joint update and the augmentations of one
    observation step (test/main.cpp:188-189) run in ONE single-CTA launch — bit-identical to the two calls,
    and within 1e-9 of the oracle."""
    import conan_slam_b200 as cs
    N = 25
    X, P, lm = helpers.synthetic_map(N, 31)
    rng = np.random.default_rng(8)
    for flags in (0, cs.FLAG_INTENDED):
        a = cs.EKF(capacity_landmarks=N + 6, device=0, flags=flags)
        b = cs.EKF(capacity_landmarks=N + 6, device=0, flags=flags)
        o = oracle_py.OracleEKF(flags)
        for f in (a, b, o):
            f.reset(X, P)
        launches = []
        for case in range(4):
            mf = [3, 0, 5, 2][case]
            mn = [2, 3, 0, 1][case]
            ids = (rng.choice(N, size=mf, replace=False) + 1).astype(np.int32)
            ZF = helpers.observe(o.X, lm, ids, rng) if mf else np.zeros((2, 0))
            ZN = np.stack([rng.uniform(200, 1500, size=mn), rng.uniform(-1.5, 1.5, size=mn)]) if mn else np.zeros((2, 0))
            l0 = cs.load_library().cslam_kernel_launches()
            a.observeStep(ZF, helpers.RE, ids, ZN, True)
            launches.append(cs.load_library().cslam_kernel_launches() - l0)
            for f in (b, o):
                f.update(ZF, helpers.RE, ids, True)
                f.augment(ZN, helpers.RE)
            assert a.n == b.n == o.n
            assert np.array_equal(a.X, b.X)
            assert np.array_equal(np.triu(a.P), np.triu(b.P))
        assert launches == [1, 1, 1, 1], launches
        iu = np.triu_indices(o.n)
        assert helpers.rel_err(a.X, o.X) < 1e-9 and helpers.rel_err(a.P[iu], o.P[iu]) < 1e-9
        assert a.sync() == b.sync()  # numerically skipped updates (slam.h:252-255) are counted identically
        a.close()
        b.close()
