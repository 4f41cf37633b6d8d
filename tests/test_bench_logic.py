"""The pure-Python bookkeeping of bench.py (CPU tier): which roofline block a timed loop gets, its arithmetic, the
config object shared with the reference arm.  No GPU, no library call — the timed dictionaries are made up."""
import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    mod.dmma_peak = lambda dev: (37.0, "fixed for the test")
    return mod


class _Ctx:
    def __init__(self, world, sharded):
        self.world, self.sharded, self.local, self.rank = world, sharded, 0, 0


def _timed(n, launches, updates, ms_per_launch, shards=1, total_ms=10.0):
    return {"cov_ms": ms_per_launch * launches, "cov_launches": launches,
            "cov_bytes": 8.0 * n * (n + 1) * launches / shards, "updates": updates, "ms": total_ms}


def test_roofline_block_follows_the_kernel_that_ran(bench):
    n, peak = 40003, 6550.7
    one = _Ctx(1, False)
    # 64-row banks: 32 updates per launch -> tensor-bound block against the DMMA peak, HBM figures beside it
    r = bench.ekf_roofline(one, n, 20000, 4, _timed(n, 2, 64, 3.6), False, False, peak, "m")
    assert r["bound"] == "tensor" and r["kernel"].startswith("k_cov_update_dmma") and r["unit"] == "TFLOP/s"
    flops = 64.0 * n * (n + 1)
    assert r["achieved"] == pytest.approx(flops / 3.6e-3 / 1e12) and r["frac"] == pytest.approx(r["achieved"] / 37.0)
    assert r["hbm_gbs_same_launch"] == pytest.approx(8.0 * n * (n + 1) / 3.6e-3 / 1e9)
    assert r["traffic"] is not None          # the committed ncu capture of that kernel at this size
    # 16-row banks: 8 updates per launch -> the TMA streaming pass, HBM-bound
    r = bench.ekf_roofline(one, n, 20000, 4, _timed(n, 10, 80, 2.14), False, False, peak, "m")
    assert r["bound"] == "hbm" and r["kernel"].startswith("k_cov_update_tma_dense")
    assert r["frac"] == pytest.approx(8.0 * n * (n + 1) / 2.14e-3 / 1e9 / peak) and 0.9 < r["frac"] < 0.92
    assert r["traffic"] is not None
    # one pass per update (flush after every scan): the FMA streaming kernel
    r = bench.ekf_roofline(one, n, 20000, 1, _timed(n, 12, 12, 1.99), False, True, peak, "m")
    assert r["bound"] == "hbm" and r["kernel"].startswith("k_cov_update (") and r["updates_per_launch"] == 1
    # joint update m = 32
    r = bench.ekf_roofline(one, n, 20000, 32, _timed(n, 6, 6, 3.4), True, False, peak, "m")
    assert r["bound"] == "tensor" and r["algorithmic_flops_per_launch"] == pytest.approx(64.0 * n * (n + 1))
    # sharded over 8 GPUs: per-GPU figures, no single-GPU ncu capture applies
    r = bench.ekf_roofline(_Ctx(8, True), n, 20000, 4, _timed(n, 16, 256, 0.40, shards=8), False, False, peak, "m")
    assert r["bound"] == "tensor" and r["traffic"] is None and r["per"] == "GPU"
    assert r["algorithmic_bytes_per_launch"] == pytest.approx(8.0 * n * (n + 1) / 8)
    # a small (eager) map keeps round 1's kernel name
    r = bench.ekf_roofline(one, 603, 300, 4, _timed(603, 10, 40, 0.01), False, False, peak, "m")
    assert r["kernel"].startswith("k_cov_update_multi")


def test_config_object_is_the_same_for_both_arms(bench):
    a = bench.ekf_config(20000, 4, 1, False)
    assert a["landmarks"] == 20000 and a["state_dim"] == 40003 and a["obs_per_step"] == 4
    assert "sequential update" in a["workload"]
    b = bench.ekf_config(20000, 32, 1, False, batch=True)
    assert "JOINT" in b["workload"] and b["obs_per_step"] == 32
    c = bench.ekf_config(20000, 4, 8, True)
    assert "8" in c["parallelism"] or "8 GPUs" in c["workload"]
