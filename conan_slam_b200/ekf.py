"""EKF-SLAM filter — host mirror of the reference's `class EKF : public Slam`.

Method names, argument meaning and index conventions follow slam/include/slam.h and
slam/src/EKF.cpp; the state (X, P) is owned by the GPU handle instead of the caller
(SURVEY.md §8b), so methods take no X/P arguments and accessors read them back.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import check, dptr, iptr


@dataclass
class Association:
    """slam.h:438-443 Association_t, plus the per-observation decisions of the gate."""
    ZF: np.ndarray   # 2 x #associated
    ZN: np.ndarray   # 2 x #new
    idf: np.ndarray  # 1-based map slots of ZF's columns
    jbest: np.ndarray = None
    is_new: np.ndarray = None
    nbest: np.ndarray = None
    outer: np.ndarray = None


def _z(Z):
    Z = np.asarray(Z, dtype=np.float64)
    if Z.size == 0:
        return np.zeros((2, 0)), np.zeros(0)
    if Z.ndim == 1:
        Z = Z.reshape(2, -1)
    assert Z.shape[0] == 2, "Z must be 2 x m (range; bearing), as in the reference"
    return Z, np.ascontiguousarray(Z.T).reshape(-1)  # column-major 2 x m == interleaved pairs


def _m2(M):
    return np.ascontiguousarray(np.asarray(M, dtype=np.float64).reshape(2, 2).T).reshape(-1)  # column-major


class EKF:
    """Drop-in for `std::shared_ptr<Slam> ekfSlam(new EKF(LM, WP))` (test/main.cpp:89).

    Config fields carry the reference's defaults (slam.h:65-103).
    """

    def __init__(self, landMarks=None, wayPoints=None, capacity_landmarks=None, device=0, flags=0, rank=0, world=1,
                 nccl_id=None):
        self._lib = _lib.load_library()
        self.mLM = None if landMarks is None else np.asarray(landMarks, dtype=np.float64)
        self.mWP = None if wayPoints is None else np.asarray(wayPoints, dtype=np.float64)
        if capacity_landmarks is None:
            capacity_landmarks = 0 if self.mLM is None else int(self.mLM.shape[1])
        self.mTABLE = np.zeros(0 if self.mLM is None else self.mLM.shape[1], dtype=np.int32)  # EKF.cpp:6
        # slam.h:65-103
        self.mVelocity = 83.33
        self.mMaxSWA = np.pi / 4.0
        self.mRateSWA = 70.0 * np.pi / 180.0
        self.mWheelBase = 73.0
        self.mDtControls = 0.01
        self.mSigmaV = 0.3
        self.mSigmaSWA = np.pi / 180.0
        self.mMaxRange = 2000.0
        self.mSigmaR = 0.1
        self.mSigmaB = np.pi / 180.0
        self.mGateReject = 50.0
        self.mGateAugment = 1000.0
        self.mSwitchHeadingKnown = True
        self.mSwitchAssociationKnown = True
        self.mSwitchBatchUpdate = True
        self.flags = flags
        h = C.c_void_p()
        self.rank, self.world = int(rank), int(world)
        if world > 1:  # row-sharded covariance: one process per GPU, SPMD calls (include/cslam.h)
            assert nccl_id is not None and len(nccl_id) == 128, "pass the 128-byte id from dist.nccl_unique_id()"
            buf = C.create_string_buffer(bytes(nccl_id), 128)
            check(self._lib.cslam_ekf_create_sharded(C.byref(h), int(capacity_landmarks), int(device), int(flags),
                                                     int(rank), int(world), C.cast(buf, C.c_void_p)),
                  "cslam_ekf_create_sharded")
            self._h = h
            import os
            if os.environ.get("CSLAM_EKF_PEER", "1") != "0":
                self._exchange_ipc(device)
        else:
            check(self._lib.cslam_ekf_create(C.byref(h), int(capacity_landmarks), int(device), int(flags)),
                  "cslam_ekf_create")
        self._h = h

    def _exchange_ipc(self, device):
        """Peer-memory column exchange: every rank exports the CUDA-IPC handles of its snapshot buffers and
        flags, torch.distributed carries them around, every rank maps its peers' (cslam_ekf_ipc_*)."""
        import torch
        import torch.distributed as dist
        mine = C.create_string_buffer(128)
        rc = self._lib.cslam_ekf_ipc_export(self._h, C.cast(mine, C.c_void_p))
        if rc != 0:  # not a lazy handle (CSLAM_LAZY=0): the columns travel by NCCL
            return
        dev = f"cuda:{device}" if dist.get_backend() == "nccl" else "cpu"
        t = torch.frombuffer(bytearray(mine.raw), dtype=torch.uint8).to(dev)
        parts = [torch.zeros(128, dtype=torch.uint8, device=dev) for _ in range(self.world)]
        dist.all_gather(parts, t)
        blob = b"".join(bytes(p.cpu().numpy().tobytes()) for p in parts)
        allb = C.create_string_buffer(blob, 128 * self.world)
        check(self._lib.cslam_ekf_ipc_import(self._h, C.cast(allb, C.c_void_p), self.world), "cslam_ekf_ipc_import")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.cslam_ekf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing ----------------------------------------------------------------
    def set_stream(self, cuda_stream):
        check(self._lib.cslam_ekf_set_stream(self._h, C.c_void_p(cuda_stream)), "cslam_ekf_set_stream")

    def sync(self):
        skipped = C.c_int(0)
        check(self._lib.cslam_ekf_sync(self._h, C.byref(skipped)), "cslam_ekf_sync")
        return skipped.value

    def flush(self):
        """Apply every pending (deferred) covariance term, asynchronously in stream order (cslam_ekf_flush)."""
        check(self._lib.cslam_ekf_flush(self._h), "cslam_ekf_flush")

    def pass_count(self):
        """(covariance passes launched since create, panel rows still pending)."""
        p, r = C.c_ulonglong(0), C.c_int(0)
        check(self._lib.cslam_ekf_pass_count(self._h, C.byref(p), C.byref(r)), "cslam_ekf_pass_count")
        return int(p.value), int(r.value)

    def profile_begin(self, max_launches=4096):
        check(self._lib.cslam_ekf_profile_begin(self._h, int(max_launches)), "cslam_ekf_profile_begin")

    def profile_end(self):
        """(summed covariance-kernel ms, launches, summed algorithmic bytes)."""
        ms, cnt, by = C.c_double(0), C.c_int(0), C.c_double(0)
        check(self._lib.cslam_ekf_profile_end(self._h, C.byref(ms), C.byref(cnt), C.byref(by)),
              "cslam_ekf_profile_end")
        return ms.value, cnt.value, by.value

    @property
    def n(self):
        return self._lib.cslam_ekf_n(self._h)

    @property
    def num_landmarks(self):
        return self._lib.cslam_ekf_num_landmarks(self._h)

    def reset(self, X, P=None):
        X = np.ascontiguousarray(X, dtype=np.float64)
        n = X.shape[0]
        pp = None
        if P is not None:
            P = np.ascontiguousarray(P, dtype=np.float64)
            assert P.shape == (n, n)
            pp = dptr(P)
        check(self._lib.cslam_ekf_reset(self._h, dptr(X), n, pp), "cslam_ekf_reset")

    def save(self, path):
        """Checkpoint X and the upper triangle of P (the reference has no persistence)."""
        check(self._lib.cslam_ekf_save(self._h, str(path).encode()), "cslam_ekf_save")

    def load(self, path):
        check(self._lib.cslam_ekf_load(self._h, str(path).encode()), "cslam_ekf_load")

    def landmark_covs(self, first=1, count=None):
        """2x2 marginal covariances of landmarks first .. first+count-1 (1-based) as (count, 2, 2)."""
        if count is None:
            count = self.num_landmarks - first + 1
        out = np.zeros((count, 3), dtype=np.float64)
        if count:
            check(self._lib.cslam_ekf_get_landmark_covs(self._h, int(first), int(count), dptr(out)),
                  "cslam_ekf_get_landmark_covs")
        return np.stack([np.stack([out[:, 0], out[:, 1]], -1), np.stack([out[:, 1], out[:, 2]], -1)], -2)

    # -- accessors (the driver owns X,P in the reference) -------------------------
    @property
    def X(self):
        n = self.n
        out = np.empty(n, dtype=np.float64)
        check(self._lib.cslam_ekf_get_state(self._h, dptr(out), n), "cslam_ekf_get_state")
        return out

    def cov_block(self, r0, c0, nr, nc):
        out = np.empty((nr, nc), dtype=np.float64)
        check(self._lib.cslam_ekf_get_cov_block(self._h, r0, c0, nr, nc, dptr(out)), "cslam_ekf_get_cov_block")
        return out

    def cov_gather(self, idx):
        """Principal sub-matrix P[idx][:, idx] (0-based state indices) without reading P back."""
        idx = np.ascontiguousarray(idx, dtype=np.int32).reshape(-1)
        k = idx.shape[0]
        out = np.empty((k, k), dtype=np.float64)
        if k:
            check(self._lib.cslam_ekf_get_cov_gather(self._h, iptr(idx), k, dptr(out)), "cslam_ekf_get_cov_gather")
        return out

    @property
    def P(self):
        n = self.n
        return self.cov_block(0, 0, n, n)

    # -- Slam interface -----------------------------------------------------------
    def predict(self, v, swa, Q, wb, dt):
        """slam.h:841-847 / EKF.cpp:406-455."""
        q = _m2(Q)
        check(self._lib.cslam_ekf_predict(self._h, float(v), float(swa), dptr(q), float(wb), float(dt)),
              "cslam_ekf_predict")

    def observeHeading(self, phi, useHeading=False):
        """slam.h:788 / EKF.cpp:328-352."""
        check(self._lib.cslam_ekf_observe_heading(self._h, float(phi), int(bool(useHeading))),
              "cslam_ekf_observe_heading")

    def update(self, Z, R, idf, batch=False):
        """slam.h:938-943 / EKF.cpp:481-496."""
        Zm, zflat = _z(Z)
        idf = np.ascontiguousarray(idf, dtype=np.int32).reshape(-1)
        m = Zm.shape[1]
        assert idf.shape[0] == m
        if m == 0:
            return
        r = _m2(R)
        if batch:
            check(self._lib.cslam_ekf_update(self._h, dptr(zflat), iptr(idf), m, dptr(r), 1), "cslam_ekf_update")
        else:
            for b in range(0, m, _lib.MAX_OBS):
                zc = np.ascontiguousarray(zflat[2 * b:2 * (b + _lib.MAX_OBS)])
                ic = np.ascontiguousarray(idf[b:b + _lib.MAX_OBS])
                check(self._lib.cslam_ekf_update(self._h, dptr(zc), iptr(ic), ic.shape[0], dptr(r), 0),
                      "cslam_ekf_update")

    def singleUpdate(self, Z, R, idf):
        self.update(Z, R, idf, batch=False)

    def batchUpdate(self, Z, R, idf):
        self.update(Z, R, idf, batch=True)

    def augment(self, Z, R):
        """slam.h:190-191 / EKF.cpp:9-26."""
        Zm, zflat = _z(Z)
        m = Zm.shape[1]
        if m == 0:
            return
        r = _m2(R)
        for b in range(0, m, _lib.MAX_OBS):
            zc = np.ascontiguousarray(zflat[2 * b:2 * (b + _lib.MAX_OBS)])
            check(self._lib.cslam_ekf_augment(self._h, dptr(zc), zc.shape[0] // 2, dptr(r)), "cslam_ekf_augment")

    def observeStep(self, ZF, R, idf, ZN, batch=True):
        """update(ZF, R, idf, batch) followed by augment(ZN, R) (test/main.cpp:188-189) as one call: on small
        maps one single-CTA launch (cslam_ekf_observe_step), bit-identical to the two calls."""
        Zf, zf = _z(ZF)
        Zn, zn = _z(ZN)
        idf = np.ascontiguousarray(idf, dtype=np.int32).reshape(-1)
        mf, mn = Zf.shape[1], Zn.shape[1]
        assert idf.shape[0] == mf
        if mf > _lib.MAX_OBS or mn > _lib.MAX_OBS:
            self.update(ZF, R, idf, batch)
            self.augment(ZN, R)
            return
        r = _m2(R)
        check(self._lib.cslam_ekf_observe_step(self._h, dptr(zf) if mf else None, iptr(idf) if mf else None, mf,
                                               dptr(zn) if mn else None, mn, dptr(r), int(bool(batch))),
              "cslam_ekf_observe_step")

    def gate(self, Z, R, gate1, gate2):
        """Per-observation gating decisions (jbest, is_new, nbest, outer)."""
        Zm, zflat = _z(Z)
        m = Zm.shape[1]
        jbest = np.zeros(m, dtype=np.int32)
        is_new = np.zeros(m, dtype=np.uint8)
        nbest = np.zeros(m, dtype=np.float64)
        outer = np.zeros(m, dtype=np.float64)
        if m:
            r = _m2(R)
            check(self._lib.cslam_ekf_gate(self._h, dptr(zflat), m, dptr(r), float(gate1), float(gate2),
                                           iptr(jbest), is_new.ctypes.data_as(_lib._u8p), dptr(nbest),
                                           dptr(outer)), "cslam_ekf_gate")
        return jbest, is_new, nbest, outer

    def controlSteps(self, v, swa, phi, useHeading, Q, wb, dt, want_trace=True):
        """k control steps (predict + observeHeading each, test/main.cpp:140-168) in one call; on small
        maps one single-CTA launch.  Returns the (k, 3) pose trace when want_trace."""
        v = np.ascontiguousarray(v, dtype=np.float64).reshape(-1)
        swa = np.ascontiguousarray(swa, dtype=np.float64).reshape(-1)
        phi = np.ascontiguousarray(phi, dtype=np.float64).reshape(-1)
        k = v.shape[0]
        assert swa.shape[0] == k and phi.shape[0] == k
        q = _m2(Q)
        trace = np.zeros((k, 3), dtype=np.float64) if want_trace else None
        check(self._lib.cslam_ekf_control_steps(self._h, k, dptr(v), dptr(swa), dptr(phi), int(bool(useHeading)),
                                                dptr(q), float(wb), float(dt), dptr(trace) if want_trace else None),
              "cslam_ekf_control_steps")
        return trace

    def scan(self, Z, R, gate1, gate2, want_indices=True):
        """dataAssociate + update(batch=False) of the associated observations as ONE asynchronous
        submission (test/main.cpp:193-195): the association indices stay on the device.  Returns
        (jbest, is_new) when want_indices (read back while the updates run), else None."""
        Zm, zflat = _z(Z)
        m = Zm.shape[1]
        if m == 0:
            return (np.zeros(0, np.int32), np.zeros(0, np.uint8)) if want_indices else None
        r = _m2(R)
        if want_indices:
            jbest = np.zeros(m, dtype=np.int32)
            is_new = np.zeros(m, dtype=np.uint8)
            check(self._lib.cslam_ekf_scan(self._h, dptr(zflat), m, dptr(r), float(gate1), float(gate2),
                                           iptr(jbest), is_new.ctypes.data_as(_lib._u8p)), "cslam_ekf_scan")
            return jbest, is_new
        check(self._lib.cslam_ekf_scan(self._h, dptr(zflat), m, dptr(r), float(gate1), float(gate2), None, None),
              "cslam_ekf_scan")
        return None

    def scan_associations(self):
        """Observations associated (= updates applied) by all fused scans so far; synchronises."""
        total = C.c_ulonglong(0)
        check(self._lib.cslam_ekf_scan_associations(self._h, C.byref(total)), "cslam_ekf_scan_associations")
        return int(total.value)

    def dataAssociate(self, Z, R, gate1, gate2):
        """slam.h:482-487 / EKF.cpp:235-326.  ZN follows Q5: empty unless FLAG Q5_RETURN_ZN."""
        Zm, _ = _z(Z)
        jbest, is_new, nbest, outer = self.gate(Zm, R, gate1, gate2)
        sel = jbest != 0
        ZF = Zm[:, sel]
        idf = jbest[sel].astype(np.int32)
        if self.flags & _lib.FLAGS["Q5_RETURN_ZN"]:
            ZN = Zm[:, is_new.astype(bool)]
        else:
            ZN = np.zeros((0, 0))
        return Association(ZF, ZN, idf, jbest, is_new, nbest, outer)

    def dataAssociateTable(self, Z, idz, table=None):
        """slam.h:454-457 / EKF.cpp:146-233 — host bookkeeping for known associations."""
        Zm, _ = _z(Z)
        idz = np.asarray(idz, dtype=np.int64).reshape(-1)
        if table is None:
            table = self.mTABLE
        zf, zn, idf, idn = [], [], [], []
        for i, ident in enumerate(idz):
            if table[ident - 1] == 0:
                zn.append(i)
                idn.append(ident)
            else:
                zf.append(i)
                idf.append(int(table[ident - 1]))
        nf = self.num_landmarks
        for k, ident in enumerate(idn):
            table[ident - 1] = nf + k + 1
        ZF = Zm[:, zf] if zf else np.zeros((0, 0))
        ZN = Zm[:, zn] if zn else np.zeros((0, 0))
        return Association(ZF, ZN, np.asarray(idf, dtype=np.int32))
