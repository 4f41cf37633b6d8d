"""Observation front-end for a simulated world (ctypes mirror of cslam_world_*): Slam::getObservations
(slam.h:575-582) with the landmarks resident on the GPU."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, dptr, iptr


class SimWorld:
    def __init__(self, landMarks, device=0):
        """landMarks: 2 x N (row 0 = x, row 1 = y), as the reference's LM (test/main.cpp:24-62)."""
        self._lib = _lib.load_library()
        lm = np.asarray(landMarks, dtype=np.float64).reshape(2, -1)
        self.num_landmarks = lm.shape[1]
        flat = np.ascontiguousarray(lm.T).reshape(-1)  # column-major 2 x N
        h = C.c_void_p()
        check(self._lib.cslam_world_create(C.byref(h), dptr(flat) if flat.size else None, self.num_landmarks,
                                           int(device)), "cslam_world_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.cslam_world_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def getObservations(self, XTrue, maxRange, max_out=None):
        """Returns (Z (2 x m), tags (m,), m_total): the visible landmarks in landmark order."""
        x = np.ascontiguousarray(XTrue, dtype=np.float64).reshape(-1)[:3].copy()
        cap = self.num_landmarks if max_out is None else int(max_out)
        Z = np.zeros(2 * max(cap, 1), dtype=np.float64)
        tags = np.zeros(max(cap, 1), dtype=np.int32)
        m = C.c_int(0)
        check(self._lib.cslam_world_observe(self._h, dptr(x), float(maxRange), cap, dptr(Z), iptr(tags), C.byref(m)),
              "cslam_world_observe")
        k = min(m.value, cap)
        return Z[:2 * k].reshape(k, 2).T.copy(), tags[:k].copy(), m.value

    def observeAndAssociate(self, XTrue, maxRange, num_map_landmarks, max_out=None):
        """getObservations + dataAssociateTable (test/main.cpp:177-186) on the device, one read-back: returns
        (ZF (2 x mf), idf (mf,), ZN (2 x mn)); the association table (mTABLE) stays on the device."""
        x = np.ascontiguousarray(XTrue, dtype=np.float64).reshape(-1)[:3].copy()
        cap = self.num_landmarks if max_out is None else int(max_out)
        ZF = np.zeros(2 * max(cap, 1), dtype=np.float64)
        ZN = np.zeros(2 * max(cap, 1), dtype=np.float64)
        idf = np.zeros(max(cap, 1), dtype=np.int32)
        mf, mn = C.c_int(0), C.c_int(0)
        check(self._lib.cslam_world_observe_associate(self._h, dptr(x), float(maxRange), int(num_map_landmarks), cap,
                                                      dptr(ZF), iptr(idf), C.byref(mf), dptr(ZN), C.byref(mn)),
              "cslam_world_observe_associate")
        kf, kn = min(mf.value, cap), min(mn.value, cap)
        return ZF[:2 * kf].reshape(kf, 2).T.copy(), idf[:kf].copy(), ZN[:2 * kn].reshape(kn, 2).T.copy()

    @property
    def table(self):
        t = np.zeros(max(self.num_landmarks, 1), dtype=np.int32)
        check(self._lib.cslam_world_get_table(self._h, iptr(t)), "cslam_world_get_table")
        return t[:self.num_landmarks]

    def reset_table(self):
        check(self._lib.cslam_world_reset_table(self._h), "cslam_world_reset_table")
