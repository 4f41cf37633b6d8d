"""Generates tests/golden/kat_r01.npz: known-answer vectors of the oracle for every hot-path operation on
small seeded inputs — predict, observeHeading, singleUpdate, batchUpdate, augment, gated association (EKF)
and predict / heading / sampleProposal / featureUpdate / resample (PF), in REF_LITERAL and INTENDED mode.
The reference ships no golden vectors; these pin OUR oracle (which is itself pinned against
the reference's own sources, tests/test_oracle_vs_ref.py) against regressions, and give the GPU parity tests
a fixed, committed target:        python tests/golden/make_kat_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers  # noqa: E402
import oracle_py  # noqa: E402


def ekf_case(flags, N=6, seed=11):
    """One fixed call sequence; returns the inputs that are not derivable and the outputs."""
    X, P, lm = helpers.synthetic_map(N, seed)
    rng = np.random.default_rng(seed)
    o = oracle_py.OracleEKF(flags)
    o.reset(X, P)
    out = {"X0": X, "P0": P, "lm": lm}
    o.predict(83.33, 0.02, helpers.QE, 73.0, 0.01)
    o.observeHeading(float(X[2]) + 1e-4, True)
    out["X_ph"], out["P_ph"] = o.X.copy(), o.P.copy()
    ids = (rng.choice(N, size=3, replace=False) + 1).astype(np.int32)
    Z = helpers.observe(o.X, lm, ids, rng)
    Zx = np.array([[7000.0], [0.5]])
    Zall = np.concatenate([Z, Zx], axis=1)
    g = o.gate(Zall, helpers.RE, 50.0, 1000.0, dense=True)
    out["ids"], out["Z"], out["Zall"] = ids, Z, Zall
    out["jbest"], out["is_new"], out["nbest"], out["outer"] = g[0], g[1], g[2], g[3]
    o.update(Z, helpers.RE, ids, False)
    out["X_seq"], out["P_seq"] = o.X.copy(), o.P.copy()
    ids2 = (rng.choice(N, size=4, replace=False) + 1).astype(np.int32)
    Z2 = helpers.observe(o.X, lm, ids2, rng)
    o.update(Z2, helpers.RE, ids2, True)
    out["ids2"], out["Z2"] = ids2, Z2
    out["X_batch"], out["P_batch"] = o.X.copy(), o.P.copy()
    Zn = np.array([[650.0, 420.0], [0.3, -1.1]])
    o.augment(Zn, helpers.RE)
    out["Zn"] = Zn
    out["X_aug"], out["P_aug"] = o.X.copy(), o.P.copy()
    return out


def pf_case(flags, npart=16, seed=5):
    rng = np.random.default_rng(seed)
    R2 = 2 * helpers.R_BASE
    o = oracle_py.OraclePF(npart, flags)
    for k in range(6):
        o.predict(83.33, 0.03, helpers.QE, 73.0, 0.01)
        o.observeHeading(0.0005 * (k + 1), True)
    xi0 = rng.normal(size=(npart, 3))
    o.samplePose(xi0)
    Z0 = np.array([[400.0, 900.0, 650.0], [0.3, -0.7, 0.05]])
    o.addOneNewFeature(Z0, R2)
    for k in range(6):
        o.predict(83.33, -0.02, helpers.QE, 73.0, 0.01)
        o.observeHeading(0.004 + 0.0005 * k, True)
    # well-conditioned pose covariances (after a pose sample P is reset to 0 and a few predicts
    # leave it close to singular — the reference inverts it as a general 3x3, PF.cpp:523-524)
    base = np.array([[4e-4, 1e-5, 1e-7], [1e-5, 5e-4, -2e-7], [1e-7, -2e-7, 3e-8]])
    covs = np.tile(base.reshape(-1), (npart, 1)) * (1.0 + 0.1 * rng.uniform(size=(npart, 1)))
    o.set_poses(o.poses, covs)
    ids = np.array([1, 3], dtype=np.int32)
    Z = np.array([[395.0, 646.0], [0.301, 0.052]])
    xi = rng.normal(size=(npart, 3))
    o.sampleProposal(Z, ids, R2, xi)
    o.featureUpdate(Z, ids, R2)
    u = rng.normal(size=npart) * 0.3
    w_before = o.weights.copy()
    w_skew = w_before * (rng.uniform(size=npart) ** 3 + 1e-3)   # uneven weights: a non-trivial resampling
    o.weights = w_skew
    keep, neff, did = o.resampleParticles(float(npart), u, True)
    return {"xi0": xi0, "Z0": Z0, "covs": covs, "ids": ids, "Z": Z, "xi": xi, "u": u, "w_before": w_before, "w_skew": w_skew, "keep": keep,
            "neff": np.float64(neff), "did": np.int32(did), "poses": o.poses.copy(), "weights": o.weights.copy(),
            "xf0": o.features(0)[0], "pf0": o.features(0)[1]}


def main():
    blob = {}
    for name, flags in (("lit", 0), ("int", oracle_py.FLAG_INTENDED)):
        for k, v in ekf_case(flags).items():
            blob[f"ekf_{name}_{k}"] = v
        for k, v in pf_case(flags).items():
            blob[f"pf_{name}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "kat_r01.npz"), **blob)
    print("wrote kat_r01.npz:", len(blob), "arrays")


if __name__ == "__main__":
    main()
