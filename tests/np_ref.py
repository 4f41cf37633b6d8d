"""Independent numpy restatement of the reference's EKF / PF arithmetic (tests only).

Written from the dense matrix formulas of slam/src/EKF.cpp, slam/src/PF.cpp and slam.h with
numpy.linalg — it shares no code with oracle/slam_oracle.hpp and guards that file against
transcription slips.  Quirk flags as in the oracle (0 = literal).
"""
import numpy as np

Q1, Q2, Q5, Q9, Q10 = 1, 2, 4, 8, 16
FLT_MIN = float(np.finfo(np.float32).tiny)


def pi2pi(a):  # slam.h:816-829
    a = np.fmod(a, 2 * np.pi)
    if a > np.pi:
        a -= 2 * np.pi
    if a < -np.pi:
        a += 2 * np.pi
    return a


def observe_model(X, idf):  # EKF.cpp:354-404
    n = X.shape[0]
    H = np.zeros((2, n))
    z = np.zeros(2)
    if n > 3:
        f = 3 + 2 * (idf - 1)
        dx, dy = X[f] - X[0], X[f + 1] - X[1]
        d2 = dx * dx + dy * dy
        d = np.sqrt(d2)
        z[:] = [d, np.arctan2(dy, dx) - X[2]]
        H[:, 0:3] = [[-dx / d, -dy / d, 0.0], [dy / d2, -dx / d2, -1.0]]
        H[:, f:f + 2] = [[dx / d, dy / d], [-dy / d2, dx / d2]]
    return z, H


def cholesky_update(X, P, V, R, H, flags=0):  # slam.h:235-266
    PHT = P @ H.T
    S = H @ PHT + R
    S = (S + S.T) * 0.5
    L = np.linalg.cholesky(S)
    Li = np.linalg.inv(L)
    G = Li.T if (flags & Q1) else Li
    W1 = PHT @ G
    W = W1 @ G.T
    return X + W @ V, P - W1 @ W1.T


def predict(X, P, v, swa, Q, wb, dt, flags=0):  # EKF.cpp:406-455
    X, P = X.copy(), P.copy()
    phi = X[2]
    s, c = np.sin(swa + phi), np.cos(swa + phi)
    Gv = np.array([[1, 0, -v * dt * s], [0, 1, v * dt * c], [0, 0, 1.0]])
    Gu = np.array([[dt * c, -v * dt * s], [dt * s, v * dt * c], [dt * np.sin(swa) / wb, v * dt * np.cos(swa) / wb]])
    P[0:3, 0:3] = Gv @ P[0:3, 0:3] @ Gv.T + Gu @ Q @ Gu.T
    n = P.shape[0]
    if n > 3:
        w = n - 3 if (flags & Q2) else n - 4
        P[0:3, 3:3 + w] = Gv @ P[0:3, 3:3 + w]
        P[3:3 + w, 0:3] = P[0:3, 3:3 + w].T
    X[0] += v * dt * c
    X[1] += v * dt * s
    X[2] = pi2pi(X[2] + v * dt * np.sin(swa) / wb)
    return X, P


def observe_heading(X, P, phi):  # EKF.cpp:328-352 + slam.h:700-725
    n = X.shape[0]
    H = np.zeros((1, n))
    H[0, 2] = 1.0
    sigma = np.float32(0.01) * np.pi / np.float32(180.0)
    R = np.array([[float(sigma) ** 2]])
    v = np.array([pi2pi(phi - X[2])])
    PHT = P @ H.T
    S = H @ PHT + R
    SI = np.linalg.inv(S)
    SI = (SI + SI.T) * 0.5
    W = PHT @ SI
    Xn = X + W @ v
    C = np.eye(n) - W @ H
    Pn = C @ P @ C.T + W @ R @ W.T
    Pn = Pn + np.eye(n) * FLT_MIN
    return Xn, Pn


def single_update(X, P, Z, R, idf, flags=0):  # EKF.cpp:457-479
    for i in range(Z.shape[1]):
        zp, H = observe_model(X, idf[i])
        V = np.array([Z[0, i] - zp[0], pi2pi(Z[1, i] - zp[1])])
        X, P = cholesky_update(X, P, V, R, H, flags)
    return X, P


def batch_update(X, P, Z, R, idf, flags=0):  # EKF.cpp:93-129
    m = Z.shape[1]
    if m == 0:
        return X, P
    n = X.shape[0]
    H = np.zeros((2 * m, n))
    V = np.zeros(2 * m)
    RR = np.zeros((2 * m, 2 * m))
    for i in range(m):
        zp, Hi = observe_model(X, idf[i])
        H[2 * i:2 * i + 2] = Hi
        V[2 * i:2 * i + 2] = [Z[0, i] - zp[0], pi2pi(Z[1, i] - zp[1])]
        RR[2 * i:2 * i + 2, 2 * i:2 * i + 2] = R
    return cholesky_update(X, P, V, RR, H, flags)


def augment(X, P, Z, R):  # EKF.cpp:9-91
    for i in range(Z.shape[1]):
        r, b = Z[0, i], Z[1, i]
        n = X.shape[0]
        s, c = np.sin(X[2] + b), np.cos(X[2] + b)
        Xn = np.concatenate([X, [X[0] + r * c, X[1] + r * s]])
        Gv = np.array([[1, 0, -r * s], [0, 1, r * c]])
        Gz = np.array([[c, -r * s], [s, r * c]])
        Pn = np.zeros((n + 2, n + 2))
        Pn[:n, :n] = P
        Pn[n:, n:] = Gv @ P[0:3, 0:3] @ Gv.T + Gz @ R @ Gz.T
        Pn[n:, 0:3] = Gv @ P[0:3, 0:3]
        Pn[0:3, n:] = Pn[n:, 0:3].T
        if n > 3:
            Pn[n:, 3:n] = Gv @ P[0:3, 3:n]
            Pn[3:n, n:] = Pn[n:, 3:n].T
        X, P = Xn, Pn
    return X, P


def compute_association(X, P, z, R, idf):  # EKF.cpp:131-144
    zp, H = observe_model(X, idf)
    V = np.array([z[0] - zp[0], pi2pi(z[1] - zp[1])])
    S = H @ P @ H.T + R
    nis = V @ np.linalg.inv(S) @ V
    nd = nis + np.log(np.linalg.det(S))
    return nis, nd


def data_associate(X, P, Z, R, gate1, gate2):  # EKF.cpp:235-326 (decisions only)
    nf = (X.shape[0] - 3) // 2
    jb, new, nbest_l, outer_l = [], [], [], []
    for i in range(Z.shape[1]):
        jbest, nbest, outer = 0, np.inf, np.inf
        for j in range(1, nf + 1):
            nis, nd = compute_association(X, P, Z[:, i], R, j)
            if nis < gate1 and nd < nbest:
                nbest, jbest = nd, j
            elif nis < outer:
                outer = nis
        jb.append(jbest)
        new.append(1 if (jbest == 0 and outer > gate2) else 0)
        nbest_l.append(nbest)
        outer_l.append(outer)
    return np.array(jb), np.array(new), np.array(nbest_l), np.array(outer_l)


# ------------------------------------------------------------------ particle filter ----
def pf_jacobians(X, xf, Pf, R):  # PF.cpp:70-135
    dx, dy = xf[0] - X[0], xf[1] - X[1]
    d2 = dx * dx + dy * dy
    d = np.sqrt(d2)
    zp = np.array([d, pi2pi(np.arctan2(dy, dx) - X[2])])
    Hv = np.array([[-dx / d, -dy / d, 0.0], [dy / d2, -dx / d2, -1.0]])
    Hf = np.array([[dx / d, dy / d], [-dy / d2, dx / d2]])
    Sf = Hf @ Pf @ Hf.T + R
    return zp, Hv, Hf, Sf


def gauss_evaluate(V, S, flags=0):  # PF.cpp:279-317
    D = V.shape[0]
    L = np.linalg.cholesky(S)
    U = L.T
    nin = np.linalg.solve(L, V) if (flags & Q9) else np.linalg.solve(U, V)
    E = -0.5 * np.sum(nin ** 2)
    C = (2 * np.pi) ** (D / 2.0) * np.prod(np.diag(U))
    return np.exp(E) / C


def pf_sample_proposal(w, X, P, XF, PF, Z, idf, R, xi, flags=0):  # PF.cpp:502-544
    X0, P0 = X.copy(), P.copy()
    X, P = X.copy(), P.copy()
    for k, ident in enumerate(idf):
        zp, Hv, Hf, Sf = pf_jacobians(X, XF[ident - 1], PF[ident - 1], R)
        Sfi = np.linalg.inv(Sf)
        V = np.array([Z[0, k] - zp[0], pi2pi(Z[1, k] - zp[1])])
        P = np.linalg.inv(Hv.T @ Sfi @ Hv + np.linalg.inv(P))
        X = X + P @ Hv.T @ Sfi @ V
    XS = np.linalg.cholesky(P) @ xi + X
    like = 1.0
    for k, ident in enumerate(idf):
        zp, Hv, Hf, Sf = pf_jacobians(XS, XF[ident - 1], PF[ident - 1], R)
        V = np.array([Z[0, k] - zp[0], pi2pi(Z[1, k] - zp[1])])
        like *= gauss_evaluate(V, Sf, flags)
    d0 = X0 - XS
    d0[2] = pi2pi(d0[2])
    d1 = X - XS
    d1[2] = pi2pi(d1[2])
    prior = gauss_evaluate(d0, P0, flags)
    prop = gauss_evaluate(d1, P, flags)
    return w * like * prior / prop, XS, np.zeros((3, 3))


def pf_feature_update(X, XF, PF, Z, idf, R, flags=0):  # PF.cpp:222-277
    XF, PF = XF.copy(), PF.copy()
    for k, ident in enumerate(idf):
        zp, Hv, Hf, Sf = pf_jacobians(X, XF[ident - 1], PF[ident - 1], R)
        V = np.array([Z[0, k] - zp[0], pi2pi(Z[1, k] - zp[1])])
        XF[ident - 1], PF[ident - 1] = cholesky_update(XF[ident - 1], PF[ident - 1], V, R, Hf, flags)
    return XF, PF
