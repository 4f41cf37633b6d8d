"""Observation front-end on the GPU (cslam_world_*) against the oracle's restatement of
Slam::getObservations (slam.h:575-582 -> :608-683 -> :339-368): identical visible sets in landmark order,
range / bearing within libm rounding."""
import numpy as np
import pytest

import oracle_py

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N", [0, 1, 30, 257, 5000])
def test_get_observations_matches_oracle(N):
    import conan_slam_b200 as cs
    rng = np.random.default_rng(N + 1)
    side = 10000.0 * np.sqrt(max(N, 1) / 30.0)
    lm = rng.uniform(-side / 2, side / 2, size=(2, N))
    w = cs.SimWorld(lm)
    for t in range(6):
        pose = np.array([rng.uniform(-side / 4, side / 4), rng.uniform(-side / 4, side / 4), rng.uniform(-np.pi, np.pi)])
        rmax = [2000.0, 500.0, 1e9][t % 3]
        Zg, tg, mg = w.getObservations(pose, rmax)
        Zo, to, mo = oracle_py.get_observations(pose, lm, rmax)
        assert mg == mo and np.array_equal(tg, to)
        assert np.all(np.diff(tg) > 0)  # landmark order, as the reference's loop emits them
        if mo:
            assert np.allclose(Zg, Zo, rtol=1e-13, atol=1e-13)
    # truncated output: the true count is still reported
    if N >= 30:
        pose = np.zeros(3)
        Zg, tg, mg = w.getObservations(pose, 1e9, max_out=4)
        Zo, to, mo = oracle_py.get_observations(pose, lm, 1e9)
        assert mg == mo and np.array_equal(tg, to[:4]) and Zg.shape == (2, 4)
    w.close()


def test_front_end_feeds_the_filter():
    """getObservations on the GPU -> dataAssociateTable -> update / augment: the reference's observation
    step (test/main.cpp:177-189) with the world on the device, against the oracle fed by its own front end."""
    import conan_slam_b200 as cs
    import helpers
    rng = np.random.default_rng(9)
    lm = rng.uniform(-1500, 1500, size=(2, 40))
    w = cs.SimWorld(lm)
    g = cs.EKF(capacity_landmarks=40, flags=oracle_py.FLAG_INTENDED)
    o = oracle_py.OracleEKF(oracle_py.FLAG_INTENDED)
    pose = np.array([10.0, -5.0, 0.3])
    X0, P0 = pose.copy(), np.diag([1.0, 1.0, 1e-4])
    g.reset(X0, P0)
    o.reset(X0, P0)
    table_g, table_o = np.zeros(40, dtype=np.int64), np.zeros(40, dtype=np.int64)
    seen = 0
    for rmax in (900.0, 1400.0, 2000.0):  # a growing sensor range: known landmarks are updated, new ones join the map
        Zg, tg, _ = w.getObservations(pose, rmax)
        Zo, to, _ = oracle_py.get_observations(pose, lm, rmax)
        assert np.array_equal(tg, to) and len(tg) > seen
        seen = len(tg)
        ag = g.dataAssociateTable(Zg, tg, table_g)
        ao = g.dataAssociateTable(Zo, to, table_o)  # host bookkeeping (same code), the oracle's observations
        g.update(ag.ZF, helpers.RE, ag.idf, False)
        g.augment(ag.ZN, helpers.RE)
        o.update(ao.ZF, helpers.RE, ao.idf, False)
        o.augment(ao.ZN, helpers.RE)
    assert g.n == o.n and helpers.rel_err(g.X, o.X) < 1e-9
    iu = np.triu_indices(o.n)
    assert helpers.rel_err(g.P[iu], o.P[iu]) < 1e-9


@pytest.mark.parametrize("N", [30, 700, 4000])
def test_device_association_table_matches_host_bookkeeping(N):
    """getObservations + dataAssociateTable (test/main.cpp:177-186, EKF.cpp:146-233) on the device
    (cslam_world_observe_associate, table resident on the GPU) against the front end + the host bookkeeping
    of the adaptor: same known / new split, same map slots, same table after every scan."""
    import conan_slam_b200 as cs
    rng = np.random.default_rng(N)
    side = 10000.0 * np.sqrt(N / 30.0)
    lm = rng.uniform(-side / 2, side / 2, size=(2, N))
    w_dev, w_host = cs.SimWorld(lm), cs.SimWorld(lm)
    table = np.zeros(N, dtype=np.int64)
    nf = 0
    for t in range(8):
        pose = np.array([rng.uniform(-side / 4, side / 4), rng.uniform(-side / 4, side / 4), rng.uniform(-np.pi, np.pi)])
        rmax = [2500.0, 4000.0, 1e9][t % 3]
        Z, tags, _ = w_host.getObservations(pose, rmax)

        book_nf = nf  # EKF.cpp:212: new landmarks receive the slots nf + 1, nf + 2, ... in list order
        zf, zn, idf, idn = [], [], [], []
        for i, ident in enumerate(tags):
            if table[ident - 1] == 0:
                zn.append(i); idn.append(ident)
            else:
                zf.append(i); idf.append(int(table[ident - 1]))
        for k, ident in enumerate(idn):
            table[ident - 1] = book_nf + k + 1
        ZFd, idfd, ZNd = w_dev.observeAndAssociate(pose, rmax, nf)
        assert np.array_equal(idfd, np.asarray(idf, dtype=np.int32))
        assert ZFd.shape[1] == len(zf) and ZNd.shape[1] == len(zn)
        if zf:
            assert np.array_equal(ZFd, Z[:, zf])
        if zn:
            assert np.array_equal(ZNd, Z[:, zn])
        nf += len(zn)
        assert np.array_equal(w_dev.table, table.astype(np.int32))
    assert nf > 0
    w_dev.reset_table()
    assert not w_dev.table.any()
    for f in (w_dev, w_host):
        f.close()
