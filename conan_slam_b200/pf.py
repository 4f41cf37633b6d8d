"""FastSLAM-style particle filter — host mirror of the reference's `class PF : public Slam`.

The reference loops `for particle in particles: pf->method(particle, ...)`
(test/main.cpp:279-286,305-309,316-327); here each method runs that loop as one CUDA
kernel over the struct-of-arrays particle set owned by the handle.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, dptr, iptr
from .ekf import Association, _m2, _z


def _draws(a, count):
    """Host numpy array -> (pointer, on_device=0); int/ctypes pointer -> device pointer."""
    if isinstance(a, (int, C.c_void_p)):
        return C.c_void_p(a if isinstance(a, int) else a.value), 1, None
    arr = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
    assert arr.shape[0] == count, f"expected {count} draws, got {arr.shape[0]}"
    return C.c_void_p(arr.ctypes.data), 0, arr


class PF:
    """Drop-in for `std::shared_ptr<Slam> pfSlam(new PF(LM, WP))` + initializeParticles(n)
    (test/main.cpp:204-208; slam.h:688 -> PF.cpp:319-341)."""

    def __init__(self, num_particles=100, capacity_landmarks=30, landMarks=None, wayPoints=None, device=0, flags=0,
                 rank=0, world=1, nccl_id=None):
        self._lib = _lib.load_library()
        self.mLM = None if landMarks is None else np.asarray(landMarks, dtype=np.float64)
        self.mWP = None if wayPoints is None else np.asarray(wayPoints, dtype=np.float64)
        self.mNumParticles = int(num_particles)
        self.mNumEffective = int(0.75 * 100)  # slam.h:92-93: computed from the DEFAULT 100, not from num_particles
        self.mSwitchResample = True
        self.mSwitchHeadingKnown = True
        self.mTABLE = np.zeros(capacity_landmarks if self.mLM is None else self.mLM.shape[1], dtype=np.int32)
        self.flags = flags
        h = C.c_void_p()
        self.rank, self.world = int(rank), int(world)
        if world > 1:
            # num_particles = LOCAL count; particles are block-partitioned over the ranks (include/cslam.h)
            assert nccl_id is not None and len(nccl_id) == 128
            buf = C.create_string_buffer(bytes(nccl_id), 128)
            check(self._lib.cslam_pf_create_sharded(C.byref(h), int(num_particles), int(capacity_landmarks),
                                                    int(device), int(flags), int(rank), int(world),
                                                    C.cast(buf, C.c_void_p)), "cslam_pf_create_sharded")
            self._h = h
            self._exchange_ipc(device)
        else:
            check(self._lib.cslam_pf_create(C.byref(h), int(num_particles), int(capacity_landmarks), int(device),
                                            int(flags)), "cslam_pf_create")
            self._h = h

    def _exchange_ipc(self, device):
        """Every rank exports its buffers' CUDA-IPC handles; torch.distributed carries them around."""
        import torch
        import torch.distributed as dist
        mine = C.create_string_buffer(640)
        check(self._lib.cslam_pf_ipc_export(self._h, C.cast(mine, C.c_void_p)), "cslam_pf_ipc_export")
        dev = f"cuda:{device}" if dist.get_backend() == "nccl" else "cpu"
        t = torch.frombuffer(bytearray(mine.raw), dtype=torch.uint8).to(dev)
        parts = [torch.zeros(640, dtype=torch.uint8, device=dev) for _ in range(self.world)]
        dist.all_gather(parts, t)
        blob = b"".join(bytes(p.cpu().numpy().tobytes()) for p in parts)
        allb = C.create_string_buffer(blob, 640 * self.world)
        check(self._lib.cslam_pf_ipc_import(self._h, C.cast(allb, C.c_void_p), self.world), "cslam_pf_ipc_import")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.cslam_pf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        check(self._lib.cslam_pf_set_stream(self._h, C.c_void_p(cuda_stream)), "cslam_pf_set_stream")

    def sync(self):
        bad = C.c_int(0)
        check(self._lib.cslam_pf_sync(self._h, C.byref(bad)), "cslam_pf_sync")
        return bad.value

    def profile_begin(self, max_resamples=1024):
        check(self._lib.cslam_pf_profile_begin(self._h, int(max_resamples)), "cslam_pf_profile_begin")

    def profile_end(self):
        """(summed gather-copy ms, resamples, summed algorithmic bytes)."""
        ms, cnt, by = C.c_double(0), C.c_int(0), C.c_double(0)
        check(self._lib.cslam_pf_profile_end(self._h, C.byref(ms), C.byref(cnt), C.byref(by)), "cslam_pf_profile_end")
        return ms.value, cnt.value, by.value

    @property
    def num_particles(self):
        return self._lib.cslam_pf_num_particles(self._h)

    @property
    def num_features(self):
        return self._lib.cslam_pf_num_features(self._h)

    # -- Slam interface (each call = the reference's per-particle loop) ---------------
    def predict(self, v, swa, Q, wb, dt):
        """slam.h:858-863 / PF.cpp:419-471."""
        q = _m2(Q)
        check(self._lib.cslam_pf_predict(self._h, float(v), float(swa), dptr(q), float(wb), float(dt)),
              "cslam_pf_predict")

    def observeHeading(self, phi, useHeading=False):
        """slam.h:796 / PF.cpp:382-417."""
        check(self._lib.cslam_pf_observe_heading(self._h, float(phi), int(bool(useHeading))),
              "cslam_pf_observe_heading")

    def save(self, path):
        """Checkpoint the whole particle set (the reference has no persistence)."""
        check(self._lib.cslam_pf_save(self._h, str(path).encode()), "cslam_pf_save")

    def load(self, path):
        check(self._lib.cslam_pf_load(self._h, str(path).encode()), "cslam_pf_load")

    def controlSteps(self, v, swa, phi, useHeading, Q, wb, dt):
        """k control steps (predict + observeHeading each) of every particle in one launch."""
        v = np.ascontiguousarray(v, dtype=np.float64).reshape(-1)
        swa = np.ascontiguousarray(swa, dtype=np.float64).reshape(-1)
        phi = np.ascontiguousarray(phi, dtype=np.float64).reshape(-1)
        assert swa.shape[0] == v.shape[0] and phi.shape[0] == v.shape[0]
        q = _m2(Q)
        check(self._lib.cslam_pf_control_steps(self._h, v.shape[0], dptr(v), dptr(swa), dptr(phi),
                                               int(bool(useHeading)), dptr(q), float(wb), float(dt)),
              "cslam_pf_control_steps")

    def sampleProposal(self, Z, idf, R, xi):
        """slam.h:881-884 / PF.cpp:502-544.  xi: [P][3] standard-normal draws."""
        Zm, zflat = _z(Z)
        idf = np.ascontiguousarray(idf, dtype=np.int32).reshape(-1)
        r = _m2(R)
        ptr, on_dev, keepalive = _draws(xi, 3 * self.num_particles)
        check(self._lib.cslam_pf_sample_proposal(self._h, dptr(zflat), iptr(idf), Zm.shape[1], dptr(r), ptr, on_dev),
              "cslam_pf_sample_proposal")
        if keepalive is not None:
            self.sync()

    def featureUpdate(self, Z, idf, R):
        """slam.h:549-552 / PF.cpp:222-277."""
        Zm, zflat = _z(Z)
        idf = np.ascontiguousarray(idf, dtype=np.int32).reshape(-1)
        r = _m2(R)
        check(self._lib.cslam_pf_feature_update(self._h, dptr(zflat), iptr(idf), Zm.shape[1], dptr(r)),
              "cslam_pf_feature_update")

    def resampleParticles(self, numEffective, u, resampleStatus=False, want_keep=True, want_neff=True):
        """slam.h:871-872 / PF.cpp:473-500 (+546-596).  u: one deviate per slot.
        Returns (keep, neff, resampled).  want_keep = want_neff = False with numEffective = inf: no read-back,
        the call does not wait for the GPU (neff is then returned as None)."""
        n = self.num_particles
        ptr, on_dev, keepalive = _draws(u, n)
        keep = np.zeros(n, dtype=np.int32) if want_keep else None
        neff = C.c_double(0)
        did = C.c_int(0)
        check(self._lib.cslam_pf_resample(self._h, ptr, on_dev, float(numEffective), int(bool(resampleStatus)),
                                          iptr(keep) if want_keep else None, C.byref(neff) if want_neff else None,
                                          C.byref(did)),
              "cslam_pf_resample")
        return keep, (neff.value if want_neff else None), bool(did.value)

    def extractFeaturesFromParticles(self):
        """slam.h:517-539: the 2 x (P * nf) concatenation of every particle's feature estimates, particle order."""
        n, nf = self.num_particles, self.num_features
        out = np.zeros((n, nf, 2), dtype=np.float64)
        if nf:
            check(self._lib.cslam_pf_get_features_all(self._h, dptr(out), None), "cslam_pf_get_features_all")
        return np.ascontiguousarray(out.reshape(n * nf, 2).T)

    def scale_weights_device(self, dev_ptr):
        """w[p] *= factor[p] with factor a DEVICE vector of num_particles doubles (cslam_pf_scale_weights)."""
        check(self._lib.cslam_pf_scale_weights(self._h, C.c_void_p(int(dev_ptr)), 1), "cslam_pf_scale_weights")

    def addOneNewFeature(self, Z, R):
        """slam.h:134 / PF.cpp:9-60 for every particle."""
        Zm, zflat = _z(Z)
        if Zm.shape[1] == 0:
            return
        r = _m2(R)
        check(self._lib.cslam_pf_add_features(self._h, dptr(zflat), Zm.shape[1], dptr(r)), "cslam_pf_add_features")

    def samplePose(self, xi):
        """test/main.cpp:319-325: X = multivariateNormalGaussianDistribution(X, P, 1); P = 0."""
        ptr, on_dev, keepalive = _draws(xi, 3 * self.num_particles)
        check(self._lib.cslam_pf_sample_pose(self._h, ptr, on_dev), "cslam_pf_sample_pose")
        if keepalive is not None:
            self.sync()

    def dataAssociateTable(self, Z, idz, table=None, nf=None):
        """Known-association bookkeeping in INTENDED form (EKF.cpp:212-226 logic; the PF
        overload PF.cpp:204-213 is broken)."""
        Zm, _ = _z(Z)
        idz = np.asarray(idz, dtype=np.int64).reshape(-1)
        if table is None:
            table = self.mTABLE
        if nf is None:
            nf = self.num_features
        zf, zn, idf, idn = [], [], [], []
        for i, ident in enumerate(idz):
            if table[ident - 1] == 0:
                zn.append(i)
                idn.append(ident)
            else:
                zf.append(i)
                idf.append(int(table[ident - 1]))
        for k, ident in enumerate(idn):
            table[ident - 1] = nf + k + 1
        ZF = Zm[:, zf] if zf else np.zeros((0, 0))
        ZN = Zm[:, zn] if zn else np.zeros((0, 0))
        return Association(ZF, ZN, np.asarray(idf, dtype=np.int32))

    def extractStatesFromParticles(self):
        """slam.h:493-511 — pose of the MINIMUM-weight particle."""
        X = np.zeros(3)
        idx = C.c_int(0)
        check(self._lib.cslam_pf_extract_state(self._h, dptr(X), C.byref(idx)), "cslam_pf_extract_state")
        return X, idx.value

    # -- accessors ----------------------------------------------------------------------
    @property
    def weights(self):
        w = np.empty(self.num_particles)
        check(self._lib.cslam_pf_get_weights(self._h, dptr(w)), "cslam_pf_get_weights")
        return w

    @weights.setter
    def weights(self, w):
        w = np.ascontiguousarray(w, dtype=np.float64)
        assert w.shape[0] == self.num_particles
        check(self._lib.cslam_pf_set_weights(self._h, dptr(w)), "cslam_pf_set_weights")

    @property
    def poses(self):
        X = np.empty((self.num_particles, 3))
        check(self._lib.cslam_pf_get_poses(self._h, dptr(X)), "cslam_pf_get_poses")
        return X

    @property
    def pose_covs(self):
        P = np.empty((self.num_particles, 3, 3))
        check(self._lib.cslam_pf_get_pose_covs(self._h, dptr(P)), "cslam_pf_get_pose_covs")
        return P

    def set_poses(self, X, Pv=None):
        X = np.ascontiguousarray(X, dtype=np.float64).reshape(self.num_particles, 3)
        pp = None
        if Pv is not None:
            Pv = np.ascontiguousarray(Pv, dtype=np.float64).reshape(self.num_particles, 9)
            pp = dptr(Pv)
        check(self._lib.cslam_pf_set_poses(self._h, dptr(X), pp), "cslam_pf_set_poses")

    def features(self, particle):
        nf = self.num_features
        XF = np.zeros((nf, 2))
        PF = np.zeros((nf, 2, 2))
        check(self._lib.cslam_pf_get_features(self._h, int(particle), dptr(XF), dptr(PF)), "cslam_pf_get_features")
        return XF, PF
