"""The algorithm behind k_gain_group_lazy (conan_slam_b200/csrc/ekf_lazy.cuh), modelled in numpy (CPU tier).

What the kernel relies on, checked here against the plain dense sequence (np_ref.single_update, a restatement of
EKF.cpp:457-479 + slam.h:235-266):
  * the g sequential, re-linearised updates of a scan depend on each other only through the marginal over
    M = pose + the scan's landmarks — a 3 + 2g system replays them and yields H_k, G_k, v_k and W1_k at the rows of M;
  * every other state row then needs only its own entries at the columns of M: W1_k(i) = P(i, M) H_k^T G_k, followed
    by P(i, M) -= W1_k(i) W1_k(M)^T — no row talks to another;
  * the covariance array may lag behind by pending rank-1 terms (P_dev = P_true + sum a a^T): entries read from a
    column snapshot of P_dev are brought up to date by subtracting the pending terms at the columns of M, while rows
    0..2 and the 2x2 diagonal blocks are kept current on the side;
  * a landmark observed twice in one scan is just two identical columns of M.
"""
import numpy as np
import pytest

import np_ref


def _spd_state(rng, N):
    n = 3 + 2 * N
    A = rng.normal(size=(n, n)) * 0.05
    P = A @ A.T + np.diag(rng.uniform(0.5, 1.5, size=n))
    X = np.zeros(n)
    X[:3] = [1.0, -2.0, 0.1]
    ang = rng.uniform(-np.pi, np.pi, size=N)
    rad = rng.uniform(50.0, 400.0, size=N)
    X[3::2] = X[0] + rad * np.cos(ang)
    X[4::2] = X[1] + rad * np.sin(ang)
    return X, P


def _observe(rng, X, idf):
    Z = np.zeros((2, len(idf)))
    for k, j in enumerate(idf):
        z, _ = np_ref.observe_model(X, j)
        Z[:, k] = [z[0] + rng.normal() * 0.1, np_ref.pi2pi(z[1] + rng.normal() * 0.01)]
    return Z


def group_update_through_marginal(X, R3, D, P_dev, pending, Z, R, idf, flags):
    """Numpy model of one k_gain_group_lazy launch.  Returns the new X, the 2g panel rows W1 (one per rank-1 term),
    and the followed R3 / D.  P_dev is only read through the column snapshot; `pending` are the rows a_t."""
    n, g = X.shape[0], len(idf)
    f = [3 + 2 * (j - 1) for j in idf]
    cols = [0, 1, 2] + [c for fk in f for c in (fk, fk + 1)]          # columns of M (duplicates allowed)
    d = len(cols)
    # column snapshot of the lagging array; rows 0..2 come from the always-current R3
    snap = np.array([P_dev[:, c] for c in cols[3:]])                   # [2g][n]
    snap[:, :3] = np.array([R3[:, c] for c in cols[3:]])
    Ac = np.array([[a[c] for c in cols] for a in pending]).reshape(len(pending), d)   # header: pending terms at M

    def load_row(i):      # P_true(i, M): pose part from R3 (current), landmark part = snapshot - pending terms
        row = np.empty(d)
        row[:3] = R3[:, i]
        row[3:] = snap[:, i]
        if i >= 3:
            for t, a in enumerate(pending):
                row[3:] -= a[i] * Ac[t, 3:]
        return row

    # ---- replay on the marginal rows
    Pm = np.array([load_row(c) for c in cols])                         # d x d
    Xm = np.array([X[c] for c in cols])
    small = []
    for k in range(g):
        sel = [0, 1, 2, 3 + 2 * k, 4 + 2 * k]
        x5, Pc = Xm[sel], Pm[np.ix_(sel, sel)]
        z, H5 = np_ref.observe_model(x5, 1)                            # 5-state system: pose + this landmark
        v = np.array([Z[0, k] - z[0], np_ref.pi2pi(Z[1, k] - z[1])])
        S = H5 @ Pc @ H5.T + R
        S = (S + S.T) * 0.5
        Li = np.linalg.inv(np.linalg.cholesky(S))
        G = Li.T if (flags & np_ref.Q1) else Li
        W1m = Pm[:, sel] @ H5.T @ G                                    # d x 2: W1 at the rows of M
        Xm = Xm + W1m @ G.T @ v
        Pm = Pm - W1m @ W1m.T
        small.append((sel, H5, G, v, W1m))
    # ---- every row on its own
    Xn, W1 = X.copy(), np.zeros((2 * g, n))
    R3n, Dn = R3.copy(), D.copy()
    for i in range(n):
        row = load_row(i)
        for k, (sel, H5, G, v, W1m) in enumerate(small):
            w1 = row[sel] @ H5.T @ G                                   # 2 values
            Xn[i] += w1 @ G.T @ v
            row = row - w1 @ W1m.T
            W1[2 * k:2 * k + 2, i] = w1
        R3n[:, i] = row[:3]                                            # followed rows 0..2 (P(i, 0..2) by symmetry)
    for l in range((n - 3) // 2):                                      # followed 2x2 diagonal blocks
        i = 3 + 2 * l
        for q in range(2 * g):
            a0, a1 = W1[q, i], W1[q, i + 1]
            Dn[l] -= np.array([[a0 * a0, a0 * a1], [a0 * a1, a1 * a1]])
    return Xn, W1, R3n, Dn


@pytest.mark.parametrize("flags", [np_ref.Q1, 0])
@pytest.mark.parametrize("idf", [[5], [2, 9], [7, 3, 11, 4], [6, 1, 6, 10], [1, 2, 3, 4, 5, 6, 7, 8]])
def test_group_replay_equals_the_dense_sequence(flags, idf):
    if flags == 0 and len(set(idf)) < len(idf):
        # the literal gain (G = L^-1) is not the Kalman gain: observing a landmark twice in one scan can
        # leave S indefinite, where the reference takes its zero-gain path — outside this model
        pytest.skip("literal gain + a landmark observed twice: S may lose definiteness")
    rng = np.random.default_rng(1000 + len(idf) + flags)
    N = 12
    X, P_true = _spd_state(rng, N)
    n = X.shape[0]
    # the covariance array lags behind by 5 pending rank-1 terms (rows >= 3; rows 0..2 and D are current)
    pending = [rng.normal(size=n) * 0.02 for _ in range(5)]
    P_dev = P_true + sum(np.outer(a, a) for a in pending)
    R3 = P_true[:3, :].copy()
    D = np.array([P_true[3 + 2 * l:5 + 2 * l, 3 + 2 * l:5 + 2 * l] for l in range(N)])
    R = np.diag([0.1 ** 2, (np.pi / 180.0) ** 2])
    idf = np.asarray(idf, dtype=int)
    Z = _observe(rng, X, idf)

    Xs, Ps = np_ref.single_update(X.copy(), P_true.copy(), Z, R, idf, flags)            # the reference's sequence
    Xg, W1, R3g, Dg = group_update_through_marginal(X, R3, D, P_dev, pending, Z, R, idf, flags)
    Pg = P_dev - sum(np.outer(a, a) for a in pending) - W1.T @ W1                         # what the next pass applies

    scale = np.max(np.abs(Ps))
    assert np.max(np.abs(Xg - Xs)) < 1e-10 * max(1.0, np.max(np.abs(Xs)))
    assert np.max(np.abs(Pg - Ps)) < 1e-10 * scale
    assert np.max(np.abs(R3g - Ps[:3, :])) < 1e-10 * scale                                # rows 0..2 stayed current
    for l in range(N):
        assert np.max(np.abs(Dg[l] - Ps[3 + 2 * l:5 + 2 * l, 3 + 2 * l:5 + 2 * l])) < 1e-10 * scale


def test_flagged_cell_protocol_model():
    """k_col_push_ll / ll_wait: a double travels as {low word, epoch, high word, epoch}; each 8-byte half carries its
    own tag, so whichever way a 16-byte store is split on its way, a reader that sees both tags equal to the epoch it
    waits for holds exactly the value that was sent for that epoch."""
    rng = np.random.default_rng(7)

    def pack(v, epoch):
        bits = np.array([v], dtype=np.float64).view(np.uint64)[0]
        return np.array([bits & 0xFFFFFFFF, epoch, bits >> 32, epoch], dtype=np.uint64)

    def ready(cell, epoch):
        return cell[1] == epoch and cell[3] == epoch

    def unpack(cell):
        return np.array([(int(cell[2]) << 32) | int(cell[0])], dtype=np.uint64).view(np.float64)[0]

    for _ in range(200):
        old, new = rng.normal() * 1e3, rng.normal() * 1e-3
        e = int(rng.integers(3, 1 << 31))
        cell_old, cell_new = pack(old, e - 2), pack(new, e)      # cells are reused two snapshots later
        assert ready(cell_new, e) and unpack(cell_new) == new
        # a reader that polls while only ONE half of the new cell has landed must not accept the mixture
        for torn in (np.concatenate([cell_new[:2], cell_old[2:]]), np.concatenate([cell_old[:2], cell_new[2:]])):
            assert not ready(torn, e)
        assert not ready(cell_old, e)
