#!/usr/bin/env python
"""Device-timed survey of every single-GPU configuration of BASELINE.json (development tool;
bench.py is the contract).  Prints one JSON object per configuration."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import conan_slam_b200 as cs  # noqa: E402

PEAK = bench.measured_peaks()[0]
RE = bench.RE


def timed(stream, fn, reps, flush=None):
    """Average ms of fn() over reps, CUDA events on `stream`; optional L2 flush before each rep
    (its time is excluded by per-rep events)."""
    tot = 0.0
    for _ in range(reps):
        if flush is not None:
            flush.add_(1.0)
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            fn()
            e1.record(stream)
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


def ekf_configs(which):
    stream = torch.cuda.Stream()
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device="cuda")  # 256 MB > L2
    for N in which:
        n = 3 + 2 * N
        ekf, lm, rng = bench.build_ekf(N, 0, cs.FLAG_INTENDED, stream=stream.cuda_stream)
        scans = bench.make_scans(lm, rng, 8, 4)
        for s in range(3):
            bench.ekf_scan(ekf, scans[s][0])
        ekf.sync()
        cov_bytes = 8.0 * n * (n + 1)
        upd_bytes = cov_bytes + 112.0 * n
        out = {"config": f"EKF N={N}", "n": n}
        Z, ids = scans[3]
        ids32 = ids.astype(np.int32)
        for name, fl in (("warm_l2", None), ("flushed_l2", flush)):
            ms = timed(stream, lambda: ekf.update(Z[:, :1], RE, ids32[:1], False), 10, fl)
            out[f"seq_update_ms_{name}"] = ms
            out[f"seq_update_frac_{name}"] = upd_bytes / (ms * 1e-3) / 1e9 / PEAK
        ms = timed(stream, lambda: bench.ekf_scan(ekf, Z), 10, None)
        out["scan4_ms"] = ms
        out["scan4_updates_per_s"] = 4.0 / (ms * 1e-3)
        ms = timed(stream, lambda: ekf.gate(Z, RE, 50.0, 1000.0), 10, flush)
        out["gate_ms_incl_d2h"] = ms
        ms = timed(stream, lambda: ekf.observeHeading(0.0, True), 5, flush)
        out["heading_ms"] = ms
        out["heading_frac"] = cov_bytes / (ms * 1e-3) / 1e9 / PEAK
        ms = timed(stream, lambda: ekf.predict(83.33, 0.01, 2 * bench.R_BASE, 73.0, 0.01), 10, flush)
        out["predict_ms"] = ms
        if N >= 32:
            near = np.argsort(np.hypot(lm[0], lm[1]))[:64]
            idb = rng.choice(near, size=32, replace=False)
            Zb = bench.range_bearing(np.zeros(3), lm[:, idb])
            ms = timed(stream, lambda: ekf.update(Zb, RE, (idb + 1).astype(np.int32), True), 5, flush)
            out["batch32_ms"] = ms
            out["batch32_joint_updates_per_s"] = 1e3 / ms
            out["batch32_obs_per_s"] = 32e3 / ms
            out["batch32_hbm_frac"] = (cov_bytes + 67 * 8.0 * n) / (ms * 1e-3) / 1e9 / PEAK
            out["batch32_fp64_tflops"] = 64.0 * n * (n + 1) / (ms * 1e-3) / 1e12
            ekf.profile_begin(16)
            ekf.update(Zb, RE, (idb + 1).astype(np.int32), True)
            pms, cnt, _ = ekf.profile_end()
            out["batch32_cov_kernel_ms"] = pms
        out["skipped"] = ekf.sync()
        print(json.dumps(out), flush=True)
        ekf.close()
        del ekf
        torch.cuda.empty_cache()


def pf_config(npart, nfeat, m_obs=4):
    stream = torch.cuda.Stream()
    sc = bench.PfScenario(npart, nfeat, m_obs, 0, seed=npart + 2, stream=stream.cuda_stream)
    pf = sc.pf
    xi = torch.randn(npart, 3, dtype=torch.float64, device="cuda")
    u = torch.randn(npart, dtype=torch.float64, device="cuda") * 0.3
    torch.cuda.synchronize()
    sc.init_map(xi.data_ptr())
    did = []
    for _ in range(2):
        did.append(sc.cycle_step(xi.data_ptr(), u.data_ptr())[2])
    pf.sync()
    ms = timed(stream, lambda: did.append(sc.cycle_step(xi.data_ptr(), u.data_ptr())[2]), 5, None)
    w = pf.weights
    bytes_per = bench.PF_CONTROLS_PER_OBS * 2 * 208 + m_obs * 96 + 2 * (nfeat * 40 + 13 * 8)
    out = {"config": f"PF {npart} particles x {nfeat} landmarks", "step_ms": ms,
           "particle_steps_per_s": npart / (ms * 1e-3), "bytes_per_particle_step": bytes_per,
           "hbm_frac": bytes_per * npart / (ms * 1e-3) / 1e9 / PEAK, "bad": pf.sync(),
           "resampled_every_cycle": bool(all(did)), "weights_finite": bool(np.all(np.isfinite(w)))}
    Z, ids = sc.observation()
    for name, fn in (("predict", lambda: pf.predict(83.33, 0.02, bench.PF_Q, 73.0, 0.01)),
                     ("heading", lambda: pf.observeHeading(sc.pose[2], True)),
                     ("resample", lambda: pf.resampleParticles(float("inf"), u.data_ptr(), True, want_keep=False))):
        out[name + "_ms"] = timed(stream, fn, 3, None)
    print(json.dumps(out), flush=True)
    pf.close()


if __name__ == "__main__":
    what = sys.argv[1:] or ["ekf", "pf"]
    if "ekf" in what:
        ekf_configs([30, 2000, 20000])
    if "pf" in what:
        pf_config(100000, 500)
        pf_config(1000000, 500)
