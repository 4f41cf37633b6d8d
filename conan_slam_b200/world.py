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
