"""conan_slam_b200 — B200-native (sm_100a) filter hot path of conan-slam.

Host-side mirror of the reference's `Slam` filter interface (slam/include/slam.h) over the
C ABI of libcslam.so (include/cslam.h).  All arithmetic runs in hand-written CUDA kernels;
there is no CPU fallback: importing works anywhere, but every compute call needs a GPU and
the in-tree built library (python -m conan_slam_b200.build).
"""
from ._lib import CslamError, FLAG_INTENDED, FLAG_REF_LITERAL, FLAGS, lib_path, load_library  # noqa: F401
from .ekf import EKF, Association  # noqa: F401
from .pf import PF  # noqa: F401
from .world import SimWorld  # noqa: F401

__all__ = ["EKF", "PF", "Association", "CslamError", "FLAGS", "FLAG_REF_LITERAL", "FLAG_INTENDED",
           "load_library", "lib_path"]
