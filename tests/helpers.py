"""Shared generators for the parity tests: seeded synthetic maps in the reference's density."""
import numpy as np

import np_ref

R_BASE = np.diag([0.1 ** 2, (np.pi / 180.0) ** 2])  # slam.h:80-81
RE = 8.0 * R_BASE                                   # test/main.cpp:128
Q_BASE = np.diag([0.3 ** 2, (np.pi / 180.0) ** 2])  # slam.h:72-73
QE = 2.0 * Q_BASE                                   # test/main.cpp:127


def synthetic_map(N, seed, pose=(10.0, -5.0, 0.3), decorrelate=3):
    """SPD joint state of N landmarks built with the filter's OWN augment:
    Pvv = diag(1, 1, (1 deg)^2), every landmark initialised from a noisy range-bearing
    observation, then a few predict/heading/update cycles (numpy restatement) so that P is a
    genuinely dense SLAM covariance.  Returns X (n), P (n x n), landmark truth (2 x N)."""
    rng = np.random.default_rng(seed)
    side = 10000.0 * np.sqrt(max(N, 1) / 30.0)
    lm = rng.uniform(-side / 2, side / 2, size=(2, N))
    X = np.array(pose, dtype=np.float64)
    P = np.diag([1.0, 1.0, (np.pi / 180.0) ** 2])
    Z = np.zeros((2, N))
    for j in range(N):
        dx, dy = lm[0, j] - X[0], lm[1, j] - X[1]
        Z[0, j] = np.hypot(dx, dy) + rng.normal() * 0.1
        Z[1, j] = np.arctan2(dy, dx) - X[2] + rng.normal() * (np.pi / 180.0)
    X, P = augment_fast(X, P, Z, RE)
    for it in range(decorrelate):
        X, P = np_ref.predict(X, P, 83.33, 0.02 * (it + 1), QE, 73.0, 0.01, flags=np_ref.Q2)
        ids = rng.choice(N, size=min(N, 3), replace=False) + 1
        Zo = observe(X, lm, ids, rng)
        X, P = np_ref.single_update(X, P, Zo, RE, ids, flags=np_ref.Q1)
        P = (P + P.T) * 0.5
    return X, P, lm


def augment_fast(X, P, Z, R):
    """EKF.cpp:28-91 for many landmarks at once (rows 0..2 are all augmentation reads)."""
    N = Z.shape[1]
    n0 = X.shape[0]
    n = n0 + 2 * N
    Xn = np.zeros(n)
    Xn[:n0] = X
    Pn = np.zeros((n, n))
    Pn[:n0, :n0] = P
    for i in range(N):
        r, b = Z[0, i], Z[1, i]
        ln = n0 + 2 * i
        s, c = np.sin(Xn[2] + b), np.cos(Xn[2] + b)
        Xn[ln:ln + 2] = [Xn[0] + r * c, Xn[1] + r * s]
        Gv = np.array([[1, 0, -r * s], [0, 1, r * c]])
        Gz = np.array([[c, -r * s], [s, r * c]])
        Pn[ln:ln + 2, ln:ln + 2] = Gv @ Pn[0:3, 0:3] @ Gv.T + Gz @ R @ Gz.T
        Pn[ln:ln + 2, 0:ln] = Gv @ Pn[0:3, 0:ln]
        Pn[0:ln, ln:ln + 2] = Pn[ln:ln + 2, 0:ln].T
    return Xn, Pn


def observe(X, lm, ids, rng, noise=True):
    """Range-bearing observations (2 x m) of the 1-based landmark ids from pose X[0:3]."""
    ids = np.asarray(ids)
    Z = np.zeros((2, ids.shape[0]))
    for k, j in enumerate(ids):
        dx, dy = lm[0, j - 1] - X[0], lm[1, j - 1] - X[1]
        Z[0, k] = np.hypot(dx, dy) + (rng.normal() * 0.1 if noise else 0.0)
        Z[1, k] = np.arctan2(dy, dx) - X[2] + (rng.normal() * np.pi / 180.0 if noise else 0.0)
    return Z


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / scale)


def rel_err_elem(a, b, floor):
    """max |a-b| / max(|b|, floor): element-wise relative error with an absolute floor."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))
