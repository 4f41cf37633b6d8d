"""ctypes face of oracle/liboracle_slam.so (the CPU restatement of the reference).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The classes mirror conan_slam_b200.EKF / PF method
for method so a parity test reads `for f in (oracle, gpu): f.predict(...)`.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle_slam.so")

FLAG_INTENDED = 0x1F

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_u8 = C.POINTER(C.c_uint8)
_LIB = None


def build_oracle(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("slam_oracle.hpp", "oracle_capi.cpp", "oracle_bench.cpp", "Makefile")]
    stale = (not os.path.exists(ORACLE_SO)) or any(os.path.getmtime(s) > os.path.getmtime(ORACLE_SO) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s", "liboracle_slam.so"])
    return ORACLE_SO


def lib():
    global _LIB
    if _LIB is None:
        build_oracle()
        L = C.CDLL(ORACLE_SO)
        L.orc_ekf_create.restype = C.c_void_p
        L.orc_ekf_create.argtypes = [C.c_uint]
        L.orc_pf_create.restype = C.c_void_p
        L.orc_pf_create.argtypes = [C.c_int, C.c_uint]
        L.orc_pi2pi.restype = C.c_double
        L.orc_pi2pi.argtypes = [C.c_double]
        L.orc_pi2pi_f.restype = C.c_float
        L.orc_pi2pi_f.argtypes = [C.c_float]
        _LIB = L
    return _LIB


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _zflat(Z):
    Z = np.asarray(Z, dtype=np.float64)
    if Z.size == 0:
        return np.zeros(0), 0
    Z = Z.reshape(2, -1)
    return np.ascontiguousarray(Z.T).reshape(-1), Z.shape[1]


def _m2(M):
    return np.ascontiguousarray(np.asarray(M, dtype=np.float64).reshape(2, 2).T).reshape(-1)


class _Backend:
    """FP64 (`orc_*`, the parity oracle) or FP32 (`orcf_*`, the reference's own precision)."""

    def __init__(self, f32=False):
        self.cdll = lib()
        self.prefix = "orcf_" if f32 else "orc_"
        self.dtype = np.float32 if f32 else np.float64
        self.ctype = C.c_float if f32 else C.c_double

    def __getattr__(self, name):
        fn = getattr(self.cdll, self.prefix + name)
        if name.endswith("_create"):
            fn.restype = C.c_void_p
        return fn

    def arr(self, a):
        return np.ascontiguousarray(a, dtype=self.dtype)

    def zflat(self, Z):
        Z = np.asarray(Z, dtype=self.dtype)
        if Z.size == 0:
            return np.zeros(0, dtype=self.dtype), 0
        Z = Z.reshape(2, -1)
        return np.ascontiguousarray(Z.T).reshape(-1), Z.shape[1]

    def m2(self, M):
        return np.ascontiguousarray(np.asarray(M, dtype=self.dtype).reshape(2, 2).T).reshape(-1)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleEKF:
    def __init__(self, flags=0, f32=False):
        self.B = _Backend(f32)
        self.L = self.B
        self.h = C.c_void_p(self.B.ekf_create(C.c_uint(flags)))
        self.flags = flags

    def __del__(self):
        try:
            self.B.ekf_destroy(self.h)
        except Exception:
            pass

    @property
    def n(self):
        return self.B.ekf_n(self.h)

    @property
    def num_landmarks(self):
        return (self.n - 3) // 2

    def reset(self, X, P=None):
        X = self.B.arr(X)
        if P is not None:
            P = self.B.arr(P)
        self.B.ekf_reset(self.h, _p(X), X.shape[0], _p(P) if P is not None else None)

    @property
    def X(self):
        out = np.empty(self.n, dtype=self.B.dtype)
        self.B.ekf_get_state(self.h, _p(out))
        return out

    @property
    def P(self):
        n = self.n
        out = np.empty((n, n), dtype=self.B.dtype)
        self.B.ekf_get_cov(self.h, _p(out))
        return out

    def predict(self, v, swa, Q, wb, dt):
        q = self.B.m2(Q)
        self.B.ekf_predict(self.h, C.c_double(v), C.c_double(swa), _p(q), C.c_double(wb), C.c_double(dt))

    def observeHeading(self, phi, useHeading=False, dense=False):
        self.B.ekf_observe_heading(self.h, C.c_double(phi), int(bool(useHeading)), int(bool(dense)))

    def update(self, Z, R, idf, batch=False):
        z, m = self.B.zflat(Z)
        if m == 0:
            return 0
        idf = np.ascontiguousarray(idf, dtype=np.int32)
        r = self.B.m2(R)
        return self.B.ekf_update(self.h, _p(z), _p(idf), m, _p(r), int(bool(batch)))

    def augment(self, Z, R):
        z, m = self.B.zflat(Z)
        if m == 0:
            return
        r = self.B.m2(R)
        self.B.ekf_augment(self.h, _p(z), m, _p(r))

    def gate(self, Z, R, gate1, gate2, dense=False):
        z, m = self.B.zflat(Z)
        r = self.B.m2(R)
        jbest = np.zeros(m, dtype=np.int32)
        is_new = np.zeros(m, dtype=np.uint8)
        nbest = np.zeros(m, dtype=self.B.dtype)
        outer = np.zeros(m, dtype=self.B.dtype)
        idf = np.zeros(max(m, 1), dtype=np.int32)
        zn = C.c_int(0)
        na = self.B.ekf_gate(self.h, _p(z), m, _p(r), C.c_double(gate1), C.c_double(gate2), int(bool(dense)),
                                 _p(jbest), is_new.ctypes.data_as(_u8), _p(nbest), _p(outer), _p(idf), C.byref(zn))
        return jbest, is_new, nbest, outer, idf[:na].copy(), zn.value

    def dataAssociateTable(self, idz, table):
        idz = np.ascontiguousarray(idz, dtype=np.int32)
        m = idz.shape[0]
        zf = np.zeros(max(m, 1), dtype=np.int32)
        zn = np.zeros(max(m, 1), dtype=np.int32)
        idf = np.zeros(max(m, 1), dtype=np.int32)
        nzf, nzn = C.c_int(0), C.c_int(0)
        self.B.ekf_table(self.h, _p(idz), m, _p(table), table.shape[0], _p(zf), _p(idf), C.byref(nzf), _p(zn),
                             C.byref(nzn))
        return zf[:nzf.value].copy(), idf[:nzf.value].copy(), zn[:nzn.value].copy()


class OraclePF:
    def __init__(self, num_particles, flags=0, f32=False):
        self.B = _Backend(f32)
        self.L = self.B
        self.np_ = int(num_particles)
        self.h = C.c_void_p(self.B.pf_create(self.np_, C.c_uint(flags)))

    def __del__(self):
        try:
            self.B.pf_destroy(self.h)
        except Exception:
            pass

    @property
    def num_particles(self):
        return self.np_

    @property
    def num_features(self):
        return self.B.pf_num_features(self.h)

    def predict(self, v, swa, Q, wb, dt):
        q = self.B.m2(Q)
        self.B.pf_predict(self.h, C.c_double(v), C.c_double(swa), _p(q), C.c_double(wb), C.c_double(dt))

    def observeHeading(self, phi, useHeading=False):
        self.B.pf_observe_heading(self.h, C.c_double(phi), int(bool(useHeading)))

    def sampleProposal(self, Z, idf, R, xi):
        z, m = self.B.zflat(Z)
        idf = np.ascontiguousarray(idf, dtype=np.int32)
        xi = self.B.arr(xi).reshape(-1)
        r = self.B.m2(R)
        self.B.pf_sample_proposal(self.h, _p(z), _p(idf), m, _p(r), _p(xi))

    def featureUpdate(self, Z, idf, R):
        z, m = self.B.zflat(Z)
        idf = np.ascontiguousarray(idf, dtype=np.int32)
        r = self.B.m2(R)
        self.B.pf_feature_update(self.h, _p(z), _p(idf), m, _p(r))

    def resampleParticles(self, numEffective, u, resampleStatus=False):
        u = self.B.arr(u)
        keep = np.zeros(self.np_, dtype=np.int32)
        neff = self.B.ctype(0)
        did = self.B.pf_resample(self.h, _p(u), C.c_double(numEffective), int(bool(resampleStatus)), _p(keep),
                                     C.byref(neff))
        return keep, neff.value, bool(did)

    def addOneNewFeature(self, Z, R):
        z, m = self.B.zflat(Z)
        if m == 0:
            return
        r = self.B.m2(R)
        self.B.pf_add_features(self.h, _p(z), m, _p(r))

    def samplePose(self, xi):
        xi = self.B.arr(xi).reshape(-1)
        self.B.pf_sample_pose(self.h, _p(xi))

    def extractStatesFromParticles(self):
        X = np.zeros(3, dtype=self.B.dtype)
        idx = self.B.pf_extract_state(self.h, _p(X))
        return X, idx

    @property
    def weights(self):
        w = np.empty(self.np_, dtype=self.B.dtype)
        self.B.pf_get_weights(self.h, _p(w))
        return w

    @weights.setter
    def weights(self, w):
        w = self.B.arr(w)
        self.B.pf_set_weights(self.h, _p(w))

    @property
    def poses(self):
        X = np.empty((self.np_, 3), dtype=self.B.dtype)
        self.B.pf_get_poses(self.h, _p(X), None)
        return X

    @property
    def pose_covs(self):
        X = np.empty((self.np_, 3), dtype=self.B.dtype)
        P = np.empty((self.np_, 3, 3), dtype=self.B.dtype)
        self.B.pf_get_poses(self.h, _p(X), _p(P))
        return P

    def set_poses(self, X, Pv=None):
        X = self.B.arr(X).reshape(-1)
        if Pv is not None:
            Pv = self.B.arr(Pv).reshape(-1)
        self.B.pf_set_poses(self.h, _p(X), _p(Pv) if Pv is not None else None)

    def features(self, particle):
        nf = self.num_features
        XF = np.zeros((nf, 2), dtype=self.B.dtype)
        PF = np.zeros((nf, 2, 2), dtype=self.B.dtype)
        self.B.pf_get_features(self.h, int(particle), _p(XF), _p(PF))
        return XF, PF


def stratified_resample(w, u, flags=0):
    """PF.cpp:546-577 alone: returns (keep, neff, cumW)."""
    L = lib()
    w = np.ascontiguousarray(w, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    n = w.shape[0]
    keep = np.zeros(n, dtype=np.int32)
    cum = np.zeros(n)
    neff = C.c_double(0)
    L.orc_stratified_resample(_d(w), _d(u), n, C.c_uint(flags), _i(keep), C.byref(neff), _d(cum))
    return keep, neff.value, cum


def get_observations(XTrue, LM, rmax, max_out=None):
    """slam.h:575-582 getObservations on a 2 x N world: (Z (2 x m), tags (m,), m_total)."""
    L = lib()
    lm = np.asarray(LM, dtype=np.float64).reshape(2, -1)
    N = lm.shape[1]
    cap = N if max_out is None else int(max_out)
    flat = np.ascontiguousarray(lm.T).reshape(-1)
    x = np.ascontiguousarray(XTrue, dtype=np.float64).reshape(-1)[:3].copy()
    Z = np.zeros(2 * max(cap, 1))
    tags = np.zeros(max(cap, 1), dtype=np.int32)
    L.orc_get_observations.restype = C.c_int
    L.orc_get_observations.argtypes = [_dp, _dp, C.c_int, C.c_double, C.c_int, _dp, _ip]
    m = L.orc_get_observations(_d(x), _d(flat), N, float(rmax), cap, _d(Z), _i(tags))
    k = min(m, cap)
    return Z[:2 * k].reshape(k, 2).T.copy(), tags[:k].copy(), m


def sim_tape(max_steps=40000, noise_seed=0, max_obs=400000):
    """Filter-independent half of test/main.cpp's loop.  Returns dict with controls [S,3]
    (vn, swan, phi_true), obs_flag [S], obs_ptr [S+1], Z [K,2], tags [K], and constants."""
    L = lib()
    controls = np.zeros((max_steps, 3))
    obs_flag = np.zeros(max_steps, dtype=np.int32)
    obs_ptr = np.zeros(max_steps + 1, dtype=np.int32)
    Z = np.zeros((max_obs, 2))
    tags = np.zeros(max_obs, dtype=np.int32)
    consts = np.zeros(8)
    L.orc_sim_tape.restype = C.c_int
    steps = L.orc_sim_tape(max_steps, C.c_ulonglong(noise_seed), _d(controls), _i(obs_flag), _i(obs_ptr), _d(Z),
                           _i(tags), max_obs, _d(consts))
    k = obs_ptr[steps]
    return dict(steps=steps, controls=controls[:steps], obs_flag=obs_flag[:steps], obs_ptr=obs_ptr[:steps + 1],
                Z=Z[:k], tags=tags[:k],
                Q=np.diag([consts[0], consts[1]]), R=np.diag([consts[2], consts[3]]), wb=consts[4], dt=consts[5])


def run_tape(filt, tape, first=0, last=None, batch=True, heading=True, table=None, on_step=None, dense_heading=None):
    """Replays test/main.cpp:132-200 (known associations) on any filter object exposing the
    reference's method names.  `table` is the feature/observation lookup table (mTABLE)."""
    QE = 2 * tape["Q"]
    RE = 8 * tape["R"]
    if table is None:
        table = np.zeros(30, dtype=np.int32)
    last = tape["steps"] if last is None else last
    for s in range(first, last):
        vn, swan, phi = tape["controls"][s]
        filt.predict(vn, swan, QE, tape["wb"], tape["dt"])
        if dense_heading is None:
            filt.observeHeading(phi, heading)
        else:
            filt.observeHeading(phi, heading, dense=dense_heading)
        if tape["obs_flag"][s]:
            a, b = tape["obs_ptr"][s], tape["obs_ptr"][s + 1]
            if b > a:
                Z = tape["Z"][a:b].T  # 2 x m
                idz = tape["tags"][a:b]
                zf, zn, idf = [], [], []
                nf = filt.num_landmarks
                new_ids = []
                for i, ident in enumerate(idz):  # EKF.cpp:146-233
                    if table[ident - 1] == 0:
                        zn.append(i)
                        new_ids.append(ident)
                    else:
                        zf.append(i)
                        idf.append(table[ident - 1])
                for k, ident in enumerate(new_ids):
                    table[ident - 1] = nf + k + 1
                if zf:
                    filt.update(Z[:, zf], RE, np.asarray(idf, dtype=np.int32), batch)
                if zn:
                    filt.augment(Z[:, zn], RE)
        if on_step is not None:
            on_step(s, filt)
    return table
