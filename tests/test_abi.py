"""CPU test: the C-ABI library builds, loads and exports every symbol include/cslam.h declares
(no compute calls — there is no GPU here), and the ctypes table matches the header."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "cslam.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cslam_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from conan_slam_b200 import _lib, build
    build.build()
    names = _declared()
    assert len(names) >= 35
    lib = C.CDLL(_lib.lib_path())
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in cslam.h but not exported: {missing}"
    unbound = [n for n in names if n not in _lib.SIGNATURES]
    assert not unbound, f"declared in cslam.h but absent from the ctypes table: {unbound}"
    extra = [n for n in _lib.SIGNATURES if n not in names]
    assert not extra, f"bound but not declared: {extra}"


def test_library_loads_and_reports_version():
    from conan_slam_b200 import _lib
    lib = _lib.load_library()
    assert lib.cslam_version() == 100
    n = C.c_int(-1)
    rc = lib.cslam_device_count(C.byref(n))
    assert rc in (0, 3)  # CSLAM_ERR_CUDA without a driver/GPU


def test_no_cpu_fallback_without_gpu(gpu_available):
    """Without a CUDA device the product refuses to run (it never routes to the oracle)."""
    if gpu_available:
        pytest.skip("GPU present")
    import conan_slam_b200 as cs
    with pytest.raises(cs.CslamError) as e:
        cs.EKF(capacity_landmarks=4)
    assert e.value.code == 3
    with pytest.raises(cs.CslamError):
        cs.PF(num_particles=8, capacity_landmarks=2)
    with pytest.raises(cs.CslamError):
        cs.SimWorld([[1.0, 2.0], [3.0, 4.0]])


def test_null_handles_are_rejected_not_dereferenced():
    """Every handle-taking compute entry returns CSLAM_ERR_BAD_ARG for a NULL handle (error convention of
    include/cslam.h: status codes, never a crash, never a silent fallback)."""
    from conan_slam_b200 import _lib
    lib = _lib.load_library()
    z = (C.c_double * 4)(1.0, 0.1, 2.0, 0.2)
    r = (C.c_double * 4)(0.01, 0.0, 0.0, 0.0003)
    ids = (C.c_int32 * 2)(1, 2)
    tot = C.c_ulonglong(0)
    m = C.c_int(0)
    calls = [
        lambda: lib.cslam_ekf_predict(None, 1.0, 0.0, r, 73.0, 0.01),
        lambda: lib.cslam_ekf_observe_heading(None, 0.0, 1),
        lambda: lib.cslam_ekf_control_steps(None, 1, z, z, z, 1, r, 73.0, 0.01, None),
        lambda: lib.cslam_ekf_scan(None, z, 2, r, 50.0, 1000.0, None, None),
        lambda: lib.cslam_ekf_update(None, z, ids, 2, r, 0),
        lambda: lib.cslam_ekf_augment(None, z, 2, r),
        lambda: lib.cslam_ekf_scan_associations(None, C.byref(tot)),
        lambda: lib.cslam_ekf_save(None, b"/tmp/x"),
        lambda: lib.cslam_pf_predict(None, 1.0, 0.0, r, 73.0, 0.01),
        lambda: lib.cslam_pf_control_steps(None, 1, z, z, z, 1, r, 73.0, 0.01),
        lambda: lib.cslam_pf_save(None, b"/tmp/x"),
        lambda: lib.cslam_world_observe(None, z, 2000.0, 2, z, ids, C.byref(m)),
    ]
    for k, call in enumerate(calls):
        assert call() == 1, f"call {k} did not report CSLAM_ERR_BAD_ARG"
    assert b"null" in lib.cslam_last_error().lower() or b"bad argument" in lib.cslam_last_error().lower()


def test_product_never_imports_oracle():
    """No include / import / dlopen of anything under oracle/ from the product tree (comments may
    cite the oracle as the parity target)."""
    pkg = os.path.join(ROOT, "conan_slam_b200")
    bad = re.compile(r"^\s*(#\s*include|import|from)\b.*oracle|CDLL\(.*oracle|dlopen\(.*oracle", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not bad.search(text), f
