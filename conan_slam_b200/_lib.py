"""ctypes binding of libcslam.so — the exact entry points include/cslam.h declares."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FLAG_REF_LITERAL = 0
FLAGS = {
    "Q1_METRIC_S": 1 << 0,
    "Q2_FULL_WIDTH": 1 << 1,
    "Q5_RETURN_ZN": 1 << 2,
    "Q9_METRIC_S": 1 << 3,
    "Q10_SEARCH": 1 << 4,
}
FLAG_INTENDED = 0x1F
MAX_OBS = 64
MAX_BATCH_OBS = 32

ERRORS = {1: "BAD_ARG", 2: "CAPACITY", 3: "CUDA", 4: "NCCL", 5: "UNSUPPORTED"}


class CslamError(RuntimeError):
    def __init__(self, code, where, msg):
        super().__init__(f"{where}: CSLAM_ERR_{ERRORS.get(code, code)}: {msg}")
        self.code = code


def lib_path():
    return os.path.join(_HERE, "lib", "libcslam.so")


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)
_vp = C.c_void_p

# name -> (restype, argtypes); mirrors include/cslam.h one to one
SIGNATURES = {
    "cslam_last_error": (C.c_char_p, []),
    "cslam_version": (C.c_int, []),
    "cslam_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "cslam_kernel_launches": (C.c_ulonglong, []),
    "cslam_ekf_profile_begin": (C.c_int, [_vp, C.c_int]),
    "cslam_ekf_profile_end": (C.c_int, [_vp, _dp, C.POINTER(C.c_int), _dp]),
    "cslam_ekf_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_uint]),
    "cslam_nccl_unique_id": (C.c_int, [_vp]),
    "cslam_ekf_create_sharded": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_uint, C.c_int, C.c_int, _vp]),
    "cslam_dmma_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "cslam_ekf_ipc_export": (C.c_int, [_vp, _vp]),
    "cslam_ekf_ipc_import": (C.c_int, [_vp, _vp, C.c_int]),
    "cslam_ekf_destroy": (C.c_int, [_vp]),
    "cslam_ekf_set_stream": (C.c_int, [_vp, _vp]),
    "cslam_ekf_sync": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "cslam_ekf_flush": (C.c_int, [_vp]),
    "cslam_ekf_pass_count": (C.c_int, [_vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_int)]),
    "cslam_ekf_n": (C.c_int, [_vp]),
    "cslam_ekf_num_landmarks": (C.c_int, [_vp]),
    "cslam_ekf_capacity": (C.c_int, [_vp]),
    "cslam_ekf_predict": (C.c_int, [_vp, C.c_double, C.c_double, _dp, C.c_double, C.c_double]),
    "cslam_ekf_observe_heading": (C.c_int, [_vp, C.c_double, C.c_int]),
    "cslam_ekf_control_steps": (C.c_int, [_vp, C.c_int, _dp, _dp, _dp, C.c_int, _dp, C.c_double, C.c_double, _dp]),
    "cslam_ekf_gate": (C.c_int, [_vp, _dp, C.c_int, _dp, C.c_double, C.c_double, _ip, _u8p, _dp, _dp]),
    "cslam_ekf_observe_step": (C.c_int, [_vp, _dp, _ip, C.c_int, _dp, C.c_int, _dp, C.c_int]),
    "cslam_ekf_scan": (C.c_int, [_vp, _dp, C.c_int, _dp, C.c_double, C.c_double, _ip, _u8p]),
    "cslam_ekf_scan_associations": (C.c_int, [_vp, C.POINTER(C.c_ulonglong)]),
    "cslam_ekf_update": (C.c_int, [_vp, _dp, _ip, C.c_int, _dp, C.c_int]),
    "cslam_ekf_augment": (C.c_int, [_vp, _dp, C.c_int, _dp]),
    "cslam_ekf_get_state": (C.c_int, [_vp, _dp, C.c_int]),
    "cslam_ekf_get_cov_block": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _dp]),
    "cslam_ekf_get_cov_gather": (C.c_int, [_vp, _ip, C.c_int, _dp]),
    "cslam_ekf_reset": (C.c_int, [_vp, _dp, C.c_int, _dp]),
    "cslam_ekf_get_landmark_covs": (C.c_int, [_vp, C.c_int, C.c_int, _dp]),
    "cslam_ekf_save": (C.c_int, [_vp, C.c_char_p]),
    "cslam_ekf_load": (C.c_int, [_vp, C.c_char_p]),
    "cslam_ekf_device_ptrs": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(C.c_size_t)]),
    "cslam_world_create": (C.c_int, [C.POINTER(_vp), _dp, C.c_int, C.c_int]),
    "cslam_world_destroy": (C.c_int, [_vp]),
    "cslam_world_observe": (C.c_int, [_vp, _dp, C.c_double, C.c_int, _dp, _ip, C.POINTER(C.c_int)]),
    "cslam_world_observe_associate": (C.c_int, [_vp, _dp, C.c_double, C.c_int, C.c_int, _dp, _ip, C.POINTER(C.c_int), _dp,
                                                C.POINTER(C.c_int)]),
    "cslam_world_get_table": (C.c_int, [_vp, _ip]),
    "cslam_world_reset_table": (C.c_int, [_vp]),
    "cslam_pf_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_uint]),
    "cslam_pf_create_sharded": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_uint, C.c_int, C.c_int, _vp]),
    "cslam_pf_ipc_export": (C.c_int, [_vp, _vp]),
    "cslam_pf_ipc_import": (C.c_int, [_vp, _vp, C.c_int]),
    "cslam_pf_destroy": (C.c_int, [_vp]),
    "cslam_pf_set_stream": (C.c_int, [_vp, _vp]),
    "cslam_pf_sync": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "cslam_pf_num_particles": (C.c_int, [_vp]),
    "cslam_pf_num_features": (C.c_int, [_vp]),
    "cslam_pf_predict": (C.c_int, [_vp, C.c_double, C.c_double, _dp, C.c_double, C.c_double]),
    "cslam_pf_observe_heading": (C.c_int, [_vp, C.c_double, C.c_int]),
    "cslam_pf_save": (C.c_int, [_vp, C.c_char_p]),
    "cslam_pf_load": (C.c_int, [_vp, C.c_char_p]),
    "cslam_pf_control_steps": (C.c_int, [_vp, C.c_int, _dp, _dp, _dp, C.c_int, _dp, C.c_double, C.c_double]),
    "cslam_pf_sample_proposal": (C.c_int, [_vp, _dp, _ip, C.c_int, _dp, _vp, C.c_int]),
    "cslam_pf_feature_update": (C.c_int, [_vp, _dp, _ip, C.c_int, _dp]),
    "cslam_pf_resample": (C.c_int, [_vp, _vp, C.c_int, C.c_double, C.c_int, _ip, _dp, C.POINTER(C.c_int)]),
    "cslam_pf_add_features": (C.c_int, [_vp, _dp, C.c_int, _dp]),
    "cslam_pf_sample_pose": (C.c_int, [_vp, _vp, C.c_int]),
    "cslam_pf_profile_begin": (C.c_int, [_vp, C.c_int]),
    "cslam_pf_profile_end": (C.c_int, [_vp, _dp, C.POINTER(C.c_int), _dp]),
    "cslam_pf_get_weights": (C.c_int, [_vp, _dp]),
    "cslam_pf_get_poses": (C.c_int, [_vp, _dp]),
    "cslam_pf_get_pose_covs": (C.c_int, [_vp, _dp]),
    "cslam_pf_get_features": (C.c_int, [_vp, C.c_int, _dp, _dp]),
    "cslam_pf_get_features_all": (C.c_int, [_vp, _dp, _dp]),
    "cslam_pf_set_weights": (C.c_int, [_vp, _dp]),
    "cslam_pf_scale_weights": (C.c_int, [_vp, _vp, C.c_int]),
    "cslam_pf_set_poses": (C.c_int, [_vp, _dp, _dp]),
    "cslam_pf_extract_state": (C.c_int, [_vp, _dp, C.POINTER(C.c_int)]),
}


def load_library():
    """Loads the in-tree libcslam.so.  Fails loudly: there is no fallback implementation."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing — build it with `python -m conan_slam_b200.build` (nvcc, sm_100a). "
            "conan_slam_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and this table ever diverge
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc, where):
    if rc != 0:
        msg = load_library().cslam_last_error()
        raise CslamError(rc, where, msg.decode() if msg else "")


def dptr(a):
    return a.ctypes.data_as(_dp)


def iptr(a):
    return a.ctypes.data_as(_ip)
