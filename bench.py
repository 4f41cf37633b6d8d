#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native conan-slam filter hot path.

Metric (BASELINE.json): EKF updates/sec at N landmarks (+ % of the HBM roofline).
One EKF update = one full observation cycle for ONE range-bearing observation: Mahalanobis
gating over all N landmarks + innovation / Kalman gain + in-place covariance update
(SURVEY.md §8d).  One benchmark "step" = one scan of m = 4 observations (cslam_ekf_scan): one gating
pass shared by the scan, 4 sequential (re-linearised) gains, and their 4 covariance updates applied in
one pass over the upper triangle — EKF.cpp:235-326 + :457-479, results bit-identical to 4 separate passes.

Default workload (N = 1 GPU): the north-star headline, a 20,000-landmark map (n = 40,003,
FP64 P = 12.8 GB, upper triangle streamed in place) — "C3-seq" of BASELINE.md.
    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
        [--batch]            C3 as worded in BASELINE.json: joint update of m = 32 observations (FP64 tensor cores)
        [--landmarks 2000]   C2          [--landmarks 60000 under torchrun]  C5
        [--workload pf]      C4: FastSLAM, 1M particles x 500 landmarks
N > 1 (torchrun, one rank per GPU): the row-sharded covariance of ONE filter (strong scaling), or
--multi replicas: independent Monte-Carlo filter replicas (no data-path collective, weak scaling).
Extra fields of the JSON line: roofline (live CUDA-event timing of the dominant kernel), cpu_baseline
(the reference's own sources on the host), drive_cycle (6 control steps + one scan, stepwise vs merged).
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R_BASE = np.diag([0.1 ** 2, (np.pi / 180.0) ** 2])
RE = 8.0 * R_BASE
GATE1, GATE2 = 50.0, 1000.0


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# Only the ONE JSON result line may reach stdout: libraries (e.g. NCCL's version banner) write to
# fd 1 directly, so fd 1 is pointed at stderr and the result goes out through a saved duplicate.
_RESULT_OUT = None


def _capture_stdout():
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


# ----------------------------------------------------------------------------- clocks ----
class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index):
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv = None
            log(f"[bench] NVML unavailable: {e}")

    def _loop(self):
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


QE_BENCH = 2.0 * np.diag([0.3 ** 2, (np.pi / 180.0) ** 2])  # test/main.cpp:127 (QE = 2Q)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, n):
    """dram bytes per launch from the committed ncu capture, if one exists for this size."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        for e in json.load(open(p)):
            if e.get("kernel") == kernel and e.get("n") == n:
                return e.get("dram_bytes_per_launch")
    return None


# ----------------------------------------------------------------------- synthetic map ----
def synth_landmarks(N, seed):
    rng = np.random.Generator(np.random.MT19937(seed))
    side = 10000.0 * np.sqrt(N / 30.0)  # the reference's density: 30 landmarks per ~(10 km)^2
    return rng.uniform(-side / 2, side / 2, size=(2, N)), rng


def range_bearing(pose, lm):
    dx, dy = lm[0] - pose[0], lm[1] - pose[1]
    return np.stack([np.hypot(dx, dy), np.arctan2(dy, dx) - pose[2]])


def build_ekf(N, device, flags, seed=None, stream=None, rank=0, world=1, nccl_id=None):
    """Synthetic N-landmark joint state built ON THE GPU by the filter's own augment kernel from
    Pvv = diag(1, 1, (1 deg)^2): P_ij = Gv_i Pvv Gv_j^T + blockdiag(Gz R Gz^T)."""
    import conan_slam_b200 as cs
    lm, rng = synth_landmarks(N, N if seed is None else seed)
    ekf = cs.EKF(capacity_landmarks=N, device=device, flags=flags, rank=rank, world=world, nccl_id=nccl_id)
    if stream is not None:
        ekf.set_stream(stream)
    pose = np.zeros(3)
    ekf.reset(pose, np.diag([1.0, 1.0, (np.pi / 180.0) ** 2]))
    Z = range_bearing(pose, lm)
    Z[0] += rng.normal(size=N) * 0.1
    Z[1] += rng.normal(size=N) * (np.pi / 180.0)
    ekf.augment(Z, RE)
    ekf.sync()
    return ekf, lm, rng


def make_scans(lm, rng, nscans, m):
    """Observation scans: m of the landmarks nearest to the (static) true pose, fresh noise."""
    d = np.hypot(lm[0], lm[1])
    near = np.argsort(d)[:max(64, m)]
    scans = []
    for s in range(nscans):
        ids = rng.choice(near, size=m, replace=False)
        Z = range_bearing(np.zeros(3), lm[:, ids])
        Z[0] += rng.normal(size=m) * 0.1
        Z[1] += rng.normal(size=m) * (np.pi / 180.0)
        scans.append((Z, ids + 1))
    return scans


def ekf_scan(ekf, Z, batch=False, want_indices=True):
    """The user-facing call sequence of one scan (test/main.cpp:193-195): gate, then update
    (batch=False: singleUpdate EKF.cpp:457-479; batch=True: one joint batchUpdate EKF.cpp:93-129)."""
    if not batch:
        # one asynchronous submission, association indices stay on the device (cslam_ekf_scan); with
        # want_indices they are read back behind the gate kernel while the updates run, without them the
        # call never waits for the GPU (the number of applied updates is read once at the end)
        r = ekf.scan(Z, RE, GATE1, GATE2, want_indices=want_indices)
        return r[0] if want_indices else None
    jbest, is_new, _, _ = ekf.gate(Z, RE, GATE1, GATE2)
    sel = jbest > 0
    ekf.update(Z[:, sel], RE, jbest[sel], batch)
    return jbest


_DMMA_PEAK = {}


def dmma_peak(device=0):
    """FP64 tensor-core peak of this GPU measured IN THIS RUN by the library's microbenchmark (cslam_dmma_peak:
    register-resident DMMA.8x8x4 chains on every SM); MEASURED_PEAKS.json carries only HBM and bf16 figures."""
    if device not in _DMMA_PEAK:
        from conan_slam_b200 import _lib
        lib = _lib.load_library()
        v = C.c_double(0.0)
        rc = lib.cslam_dmma_peak(int(device), C.byref(v))
        _DMMA_PEAK[device] = (v.value, "measured in this run (cslam_dmma_peak: register-resident DMMA.8x8x4 chains)") \
            if rc == 0 and v.value > 0 else (37.09, "round-1 measurement (profiles/dmma_bench_r01.txt)")
    return _DMMA_PEAK[device]


# ------------------------------------------------------------------------ particle filter ----
PF_Q = 2.0 * np.diag([0.3 ** 2, (np.pi / 180.0) ** 2])  # test/main.cpp:244
PF_R = 2.0 * R_BASE                                     # test/main.cpp:245
PF_CONTROLS_PER_OBS = 6  # mDtObserve / mDtControls = 5.058 -> every 6th control step (test/main.cpp:289-290)


class PfScenario:
    """Synthetic FastSLAM workload: P particles x Nf landmarks, m_obs known associations per
    observation cycle, resampling forced every cycle.  The vehicle follows the reference's
    deterministic motion model on the host (nominal pose) so that observations stay consistent
    with the particle cloud without any device read-back inside the timed loop."""

    def __init__(self, npart, nfeat, m_obs, device, seed, stream=None, rank=0, world=1, nccl_id=None):
        import conan_slam_b200 as cs
        self.cs = cs
        self.npart, self.nfeat, self.m_obs = npart, nfeat, m_obs
        self.rng = np.random.Generator(np.random.MT19937(seed))
        kw = {}
        if world > 1:
            kw = dict(rank=rank, world=world, nccl_id=nccl_id)
        self.pf = cs.PF(num_particles=npart, capacity_landmarks=nfeat, device=device, flags=cs.FLAG_INTENDED, **kw)
        if stream is not None:
            self.pf.set_stream(stream)
        self.pose = np.zeros(3)
        self.v, self.wb, self.dt = 83.33, 73.0, 0.01
        self.cycle = 0
        # landmarks ahead of the vehicle, 200 m .. 1900 m
        rng_l = np.random.Generator(np.random.MT19937(12345))
        ang = rng_l.uniform(-1.2, 1.2, size=nfeat)
        rad = rng_l.uniform(200.0, 1900.0, size=nfeat)
        self.lm = np.stack([rad * np.cos(ang) + 150.0, rad * np.sin(ang)])

    def controls(self):
        """The 6 control steps between two observations (predict + observeHeading each, test/main.cpp:279-286)
        handed over in one call: the controls and the measured heading do not depend on the filter."""
        swas, phis = [], []
        for c in range(PF_CONTROLS_PER_OBS):
            swa = 0.02 * np.sin(0.3 * (self.cycle * PF_CONTROLS_PER_OBS + c))
            # nominal pose: slam.h:952-966 vehicleModel
            x, y, phi = self.pose
            self.pose = np.array([x + self.v * self.dt * np.cos(swa + phi), y + self.v * self.dt * np.sin(swa + phi),
                                  phi + self.v * self.dt * np.sin(swa) / self.wb])
            swas.append(swa)
            phis.append(self.pose[2])
        self.pf.controlSteps(np.full(PF_CONTROLS_PER_OBS, self.v), swas, phis, True, PF_Q, self.wb, self.dt)

    def init_map(self, xi_dev_ptr):
        """6 control steps, sample the pose (test/main.cpp:319-325), initialise every landmark."""
        self.controls()
        self.pf.samplePose(xi_dev_ptr)
        Z0 = range_bearing(self.pose, self.lm)
        for b in range(0, self.nfeat, 64):
            self.pf.addOneNewFeature(Z0[:, b:b + 64], PF_R)

    def observation(self):
        ids = self.rng.choice(self.nfeat, size=self.m_obs, replace=False)
        Z = range_bearing(self.pose, self.lm[:, ids])
        Z[0] += self.rng.normal(size=self.m_obs) * 0.1
        Z[1] += self.rng.normal(size=self.m_obs) * (np.pi / 180.0)
        return Z, (ids + 1).astype(np.int32)

    def cycle_step(self, xi_ptr, u_ptr, want_keep=False):
        """One observation cycle per particle = one 'particle-step' of the metric."""
        self.controls()
        Z, ids = self.observation()
        self.pf.sampleProposal(Z, ids, PF_R, xi_ptr)
        self.pf.featureUpdate(Z, ids, PF_R)
        out = self.pf.resampleParticles(float("inf"), u_ptr, True, want_keep=want_keep)
        self.cycle += 1
        return out


# --------------------------------------------------------------------------- CPU baseline ----
def oracle_lib():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_py  # bench.py's cpu_baseline / --impl reference legs may execute oracle/
    L = oracle_py.lib()
    L.orc_bench_dense_update_slab.restype = C.c_double
    L.orc_bench_dense_update_slab.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.orc_bench_dense_gate_pair_slab.restype = C.c_double
    L.orc_bench_dense_gate_pair_slab.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    L.orc_bench_pf_step.restype = C.c_double
    L.orc_bench_pf_step.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint]
    return L


def cpu_ekf_update_rate(N, threads, slab_cols, reps=2):
    """Reference algorithm on the host cores, one sequential EKF update on an N-landmark map:
    dense choleskyUpdate (slam.h:235-266) + dense O(n^2)-per-pair gating (EKF.cpp:131-144) over
    all N landmarks.  Timed on a slab of `slab_cols` of the n columns and scaled by n/slab."""
    L = oracle_lib()
    n = 3 + 2 * N
    c = min(n, slab_cols)
    t_upd = L.orc_bench_dense_update_slab(n, c, 2, threads, reps) * n / c
    t_pair = L.orc_bench_dense_gate_pair_slab(n, c, threads, reps) * n / c
    t_full = t_upd + N * t_pair
    return {
        "value": 1.0 / t_full,
        "unit": "updates/s",
        "cores": threads,
        "kind": "port",
        "sample": (f"oracle dense choleskyUpdate + dense computeAssociation on a {c}-column slab of the "
                   f"{n}x{n} FP64 covariance, scaled by n/c; gating = N x per-pair cost (extrapolated)"),
        "update_only_updates_per_s": 1.0 / t_upd,
        "seconds_per_update_cov": t_upd,
        "seconds_per_gate_pair": t_pair,
    }


def cpu_pf_rate(nfeat, m_obs, threads, particles=2000):
    L = oracle_lib()
    t = L.orc_bench_pf_step(particles, nfeat, m_obs, threads, 0x1F)
    return {"value": 1.0 / t, "unit": "particle-steps/s", "cores": threads, "kind": "port",
            "sample": f"oracle AoS particle step (6 predict+heading, sampleProposal, featureUpdate, resample with deep "
                      f"copy) on {particles} particles x {nfeat} landmarks, per-particle cost"}


REF_SO = os.path.join(ROOT, "oracle", "_ref", "libconanslam_ref.so")


def ref_ekf_update_rate(N, budget_s=20.0):
    """The reference's OWN code (slam/src/EKF.cpp + slam.h compiled unmodified by oracle/Makefile `_ref`,
    FP32, single-threaded as the reference is) timed on a bounded sample: one EKF::update (singleUpdate ->
    Slam::choleskyUpdate, dense P*H^T and the n x n temporary) and a few EKF::computeAssociation pairs
    on an N_s-landmark map, scaled to N landmarks by the dense algorithm's own complexity — (n/n_s)^2
    per update and per gate pair, N pairs per observation (EKF.cpp:257-284).  None if oracle/_ref is
    not built (the GPU box only has what the snapshot carried)."""
    if not os.path.exists(REF_SO):
        return None
    L = C.CDLL(REF_SO)
    L.ref_ekf_create.restype = C.c_void_p
    fp, ip, vp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_void_p
    L.ref_ekf_reset.argtypes = [vp, fp, C.c_int, fp]
    L.ref_ekf_update.argtypes = [vp, fp, ip, C.c_int, fp, C.c_int]
    L.ref_ekf_compute_association.argtypes = [vp, fp, fp, C.c_int, fp, fp]
    L.ref_ekf_destroy.argtypes = [vp]
    n = 3 + 2 * N
    Rf = np.ascontiguousarray(RE, dtype=np.float32)

    def sample(Ns, pairs):
        ns = 3 + 2 * Ns
        lm, rng = synth_landmarks(Ns, Ns)
        X = np.concatenate([np.zeros(3), lm.T.reshape(-1)]).astype(np.float32)
        P = np.full((ns, ns), 1e-3, dtype=np.float32)
        P[np.arange(ns), np.arange(ns)] = 1.0
        h = C.c_void_p(L.ref_ekf_create())
        L.ref_ekf_reset(h, X.ctypes.data_as(fp), ns, P.ctypes.data_as(fp))
        del P
        j = int(np.argmin(np.hypot(lm[0], lm[1]))) + 1
        z = range_bearing(np.zeros(3), lm[:, j - 1:j]).T.reshape(-1).astype(np.float32)
        idf = np.array([j], dtype=np.int32)
        t0 = time.perf_counter()
        L.ref_ekf_update(h, z.ctypes.data_as(fp), idf.ctypes.data_as(ip), 1, Rf.ctypes.data_as(fp), 0)
        t_upd = time.perf_counter() - t0
        nis, nd = C.c_float(), C.c_float()
        t0 = time.perf_counter()
        for k in range(pairs):
            L.ref_ekf_compute_association(h, z.ctypes.data_as(fp), Rf.ctypes.data_as(fp), 1 + (j + k) % Ns,
                                          C.byref(nis), C.byref(nd))
        t_pair = (time.perf_counter() - t0) / pairs
        L.ref_ekf_destroy(h)
        return ns, t_upd, t_pair

    ns, t_upd, t_pair = sample(500, 2)                      # probe
    per_n2 = (t_upd + 3 * t_pair) / (ns * ns)
    Ns = int(min(N, max(500, (np.sqrt(budget_s / max(per_n2, 1e-12)) - 3) / 2), 8000))
    ns, t_upd, t_pair = sample(Ns, 3)
    scale = (n / ns) ** 2
    t_full = scale * t_upd + N * scale * t_pair
    return {
        "value": 1.0 / t_full, "unit": "updates/s", "cores": 1, "kind": "reference",
        "sample": (f"the reference's own EKF::update + EKF::computeAssociation (oracle/_ref: slam/src/EKF.cpp, slam.h "
                   f"compiled unmodified, FP32, single-threaded) on a {Ns}-landmark map (n_s={ns}), scaled by "
                   f"(n/n_s)^2 = {scale:.1f} per dense update and per gate pair, {N} pairs per observation (extrapolated)"),
        "update_only_updates_per_s": 1.0 / (scale * t_upd),
        "seconds_per_update_cov": scale * t_upd, "seconds_per_gate_pair": scale * t_pair,
    }


def pf_case(args, steps, warmup, spread="balanced", with_cpu=True, quiet=False):
    """C4 of BASELINE.json — FastSLAM, P particles x Nf landmarks, m_obs known associations per observation
    cycle, resampling every cycle; particles split over the GPUs.  Returns the result dict (rank 0) or None.
    spread: "balanced" = weights as the filter produces them; "realistic" = log-normal weight spread injected
    before every resampling (neff ~ 0.1-0.3 P, survivors cross rank boundaries); "adversarial" = all the
    weight mass on rank 0's particles (every other rank fetches ALL its survivors over NVLink)."""
    import torch
    import torch.distributed as dist

    import conan_slam_b200 as cs
    from conan_slam_b200 import _lib
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    lib = _lib.load_library()
    P, nfeat, m_obs = args.particles, args.pf_landmarks, args.pf_obs
    assert P % (32 * world) == 0
    Pl = P // world
    stream = torch.cuda.Stream(device=local)
    nid = None
    if world > 1:
        from conan_slam_b200 import dist as cdist
        nid = cdist.nccl_unique_id(device=f"cuda:{local}")
    sc = PfScenario(Pl, nfeat, m_obs, local, seed=P + 2, stream=stream.cuda_stream, rank=rank, world=world, nccl_id=nid)
    pf = sc.pf
    gen = torch.Generator(device=f"cuda:{local}")
    gen.manual_seed(1234 + rank)
    xi = torch.randn(Pl, 3, dtype=torch.float64, device=f"cuda:{local}", generator=gen)
    u = torch.randn(Pl, dtype=torch.float64, device=f"cuda:{local}", generator=gen) * 0.3
    xi_h = torch.empty(Pl, 3, dtype=torch.float64).pin_memory()
    u_h = torch.empty(Pl, dtype=torch.float64).pin_memory()
    xi_h.copy_(xi)
    u_h.copy_(u)
    # weight modulation for the non-balanced cases (applied on the device right before resampling)
    wmod = None
    if spread == "realistic":
        wmod = torch.exp(1.6 * torch.randn(Pl, dtype=torch.float64, device=f"cuda:{local}", generator=gen))
    elif spread == "adversarial":
        wmod = torch.full((Pl,), 1.0 if rank == 0 else 1e-300, dtype=torch.float64, device=f"cuda:{local}")
    torch.cuda.synchronize()
    sc.init_map(xi.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def cycle(want_keep=False, host_draws=False):
        sc.controls()
        Z, ids = sc.observation()
        pf.sampleProposal(Z, ids, PF_R, xi_h.numpy() if host_draws else xi.data_ptr())
        pf.featureUpdate(Z, ids, PF_R)
        if wmod is not None:
            pf.scale_weights_device(wmod.data_ptr())
        out = pf.resampleParticles(float("inf"), u_h.numpy() if host_draws else u.data_ptr(), True, want_keep=want_keep,
                                   want_neff=want_keep)
        sc.cycle += 1
        return out

    for _ in range(warmup):
        cycle()
    pf.sync()
    sampler = ClockSampler(local)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = lib.cslam_kernel_launches()
    pf.profile_begin(steps + 4)
    sampler.start()
    did = []
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(steps):
            did.append(cycle()[2])
        ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    g_ms, g_cnt, g_bytes = pf.profile_end()
    launches = lib.cslam_kernel_launches() - launches0
    ms = ev0.elapsed_time(ev1)
    # end to end: the step's random draws come from pinned HOST memory, the result (pose of the extracted
    # particle) goes back to the host every step
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(steps):
            cycle(host_draws=True)
            Xe, _ = pf.extractStatesFromParticles()  # D2H of the step's result
        e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    w = pf.weights
    ok = bool(all(did)) and bool(np.all(np.isfinite(w))) and pf.sync() == 0
    # one extra untimed cycle with the indices read back: how many survivors cross rank boundaries
    keep, neff_last, _ = cycle(want_keep=True)
    remote_frac = float(np.mean((keep // Pl) != rank))
    distinct = int(np.unique(keep).shape[0])
    distinct_remote = int(np.unique(keep[(keep // Pl) != rank]).shape[0])  # what actually crosses NVLink (duplicates coalesce)
    if not quiet:
        log(f"[bench r{rank}] PF {spread}: neff={neff_last:.1f} of {P}, remote survivors {100 * remote_frac:.1f} %, "
            f"{distinct} distinct sources for {Pl} slots")
    remote_all = remote_frac
    if world > 1:
        t = torch.tensor([ms, e2e_ms], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])
        c = torch.tensor([launches, remote_frac, distinct_remote], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        launches = int(c[0])
        remote_all = float(c[1]) / world
        distinct_remote = int(c[2])
    pf.close()
    del xi, u, wmod
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    peak, peak_src = measured_peaks()
    value = P * steps / (ms * 1e-3)
    bytes_particle = 2 * (nfeat * 40 + 13 * 8)
    bytes_step = PF_CONTROLS_PER_OBS * 2 * 208 + m_obs * 96 + bytes_particle
    ach = (g_bytes / g_cnt) / (g_ms / g_cnt * 1e-3) / 1e9 if g_cnt else 0.0
    out = {
        "metric": "PF particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"particle filter (FastSLAM) {P} particles x {nfeat} landmarks, {m_obs} known "
                               f"associations per observation cycle (6 predict+heading, sampleProposal, "
                               f"featureUpdate), resampling every cycle, particles split over {world} GPU(s)",
                   "particles": P, "landmarks": nfeat, "obs_per_step": m_obs, "mode": "INTENDED",
                   "weight_spread": spread,
                   "l2": f"inputs larger than L2 ({Pl * nfeat * 40 / 1e9:.1f} GB of particle state per GPU)",
                   "parallelism": "particles block-partitioned; one all-gather of per-rank scan totals + cumulative "
                                  "weights; survivors read from peers over NVLink inside the gather kernel"
                   if world > 1 else "single GPU"},
        "clocks": clocks,
        "e2e": {"value": P * steps / (e2e_ms * 1e-3), "unit": "particle-steps/s",
                "h2d_bytes_per_step": int(Pl * 4 * 8 + m_obs * 28 + 6 * 40), "d2h_bytes_per_step": 48},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k_gather_rows (PF.cpp:494-498 survivor copy)", "achieved": ach,
                     "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": peak_src, "traffic": None,
                     "algorithmic_bytes_per_launch": g_bytes / max(1, g_cnt), "launches_timed": g_cnt,
                     "avg_launch_ms": g_ms / max(1, g_cnt), "per": "GPU",
                     "whole_step_frac": bytes_step * Pl / (ms / steps * 1e-3) / 1e9 / peak},
        "valid": ok, "neff": neff_last, "remote_survivor_frac_rank0": remote_frac,
        "remote_survivor_frac_all_ranks": remote_all,
        "remote_slots_per_step_all_gpus": remote_all * P,
        "nvlink_bytes_per_step_all_gpus": distinct_remote * bytes_particle / 2,  # distinct remote sources x bytes read
        "distinct_sources_rank0": distinct,
    }
    if with_cpu and world == 1:
        out["cpu_baseline"] = cpu_pf_rate(nfeat, m_obs, 1)
        out["cpu_baseline"]["all_cores"] = cpu_pf_rate(nfeat, m_obs, os.cpu_count() or 1, particles=8000)
    return out


def run_pf(args):
    """--workload pf: the PF half of the metric as the whole run."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = pf_case(args, args.steps, args.warmup, spread=args.pf_spread, with_cpu=not args.no_cpu_baseline)
    if out is not None:
        emit(out)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------- main ----
def run_reference(args):
    """--impl reference: the reference's own (dense, CPU) algorithm for the same workload, all
    host threads, each step a bounded slab sample.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N = args.landmarks
    threads = os.cpu_count() or 1
    n = 3 + 2 * N
    # size the slab so that warmup + steps finish within ~2 minutes
    L = oracle_lib()
    probe_c = 256
    t_probe = L.orc_bench_dense_update_slab(n, probe_c, 2, threads, 1) + \
        L.orc_bench_dense_gate_pair_slab(n, probe_c, threads, 1)
    budget = 100.0 / max(1, args.steps + args.warmup)
    c = int(max(128, min(n, probe_c * budget / max(t_probe, 1e-6))))
    c = min(c, int(2e9 // (8 * n)))  # <= 2 GB per slab buffer
    rates = []
    use_ref = os.path.exists(REF_SO)
    for s in range(args.warmup + args.steps):
        # the reference's own compiled sources when oracle/_ref travelled with the snapshot, else the port
        r = ref_ekf_update_rate(N, budget_s=60.0 / max(1, args.steps + args.warmup)) if use_ref else \
            cpu_ekf_update_rate(N, threads, c, reps=1)
        if s >= args.warmup:
            rates.append(r)
    t_full = float(np.mean([1.0 / r["value"] for r in rates]))
    val = 1.0 / t_full
    base = rates[-1]
    base["value"] = val
    if use_ref:
        port = cpu_ekf_update_rate(N, threads, c, reps=1)
        base["oracle_port_all_cores"] = {"cores": threads, "value": port["value"], "kind": "port",
                                         "update_only_updates_per_s": port["update_only_updates_per_s"]}
    else:
        base["cores"] = threads
    out = {
        "impl": "reference", "metric": "EKF updates/sec", "value": val, "unit": "updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 4 * t_full * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if use_ref else "f64", "data": "synthetic",
        "extrapolated": True,
        "extrapolation": ("the value is EXTRAPOLATED: the reference's dense O(n^2)-per-update and O(N n^2)-per-scan "
                          "code is timed on a bounded smaller map and scaled by its own complexity (see "
                          "cpu_baseline.sample); a full-size step would take hours"),
        # the SAME config object as our arm prints for this launch (the driver compares the two lines); what the
        # reference arm actually ran is described in `reference_arm` and `cpu_baseline.sample`
        "config": ekf_config(N, 4, args.gpus, args.gpus > 1 and args.multi == "sharded"),
        "reference_arm": {"implementation": "the reference's own EKF.cpp / slam.h (oracle/_ref), dense CPU algorithm" if use_ref
                          else "oracle port of the reference's dense algorithm",
                          "arithmetic": "FP32 (as the reference)" if use_ref else "FP64", "host_threads": 1 if use_ref else threads},
        "cpu_baseline": base,
        "e2e": {"value": val, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


class Ctx:
    """Process-level context of one bench run (one rank)."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        import conan_slam_b200 as cs
        from conan_slam_b200 import _lib
        self.torch, self.dist, self.cs = torch, dist, cs
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.lib = _lib.load_library()
        self.sharded = self.world > 1 and args.multi == "sharded"
        self.dev = f"cuda:{self.local}"

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def allmax(self, *vals):
        if self.world == 1:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def allsum(self, *vals):
        if self.world == 1:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(x) for x in t]


class EkfBench:
    """One synthetic N-landmark EKF-SLAM filter on this rank's GPU (or its shard of it) + the timed loops."""

    def __init__(self, ctx, N, nscans):
        cs, torch = ctx.cs, ctx.torch
        self.ctx, self.N, self.n = ctx, N, 3 + 2 * N
        # high priority: the per-scan chain must be dispatched ahead of the pending CTAs of a running covariance pass
        self.stream = torch.cuda.Stream(device=ctx.local, priority=-1)
        t0 = time.time()
        if ctx.sharded:
            from conan_slam_b200 import dist as cdist
            nid = cdist.nccl_unique_id(device=ctx.dev)
            # one filter, covariance row-sharded over the ranks: identical inputs on every rank (SPMD)
            self.ekf, self.lm, self.rng = build_ekf(N, ctx.local, cs.FLAG_INTENDED, seed=N, stream=self.stream.cuda_stream,
                                                    rank=ctx.rank, world=ctx.world, nccl_id=nid)
        else:
            self.ekf, self.lm, self.rng = build_ekf(N, ctx.local, cs.FLAG_INTENDED, seed=N + 1000 * ctx.rank,
                                                    stream=self.stream.cuda_stream)
        log(f"[bench r{ctx.rank}] built {N}-landmark map (n={self.n}, P={8.0 * self.n * self.n / 1e9:.2f} GB) "
            f"in {time.time() - t0:.1f}s")
        self.nscans = nscans
        self.scan_cache = {}

    def scans(self, m):
        if m not in self.scan_cache:
            self.scan_cache[m] = make_scans(self.lm, self.rng, self.nscans, m)
        return self.scan_cache[m]

    def close(self):
        self.ekf.close()
        self.ctx.torch.cuda.empty_cache()

    # ---- parity against the CPU oracle on the marginal of the observed + a sample of other landmarks ----
    def parity(self, m, batch, nscan=3):
        """The filter restricted to an index set I that contains the pose and every observed landmark evolves
        exactly like the full filter (every update touches P(i,j) through rows/columns of I only), so the CPU
        oracle can replay the same calls on the |I|-dimensional marginal: state and covariance within 1e-9
        relative, association indices exactly (north_star).  Test infrastructure used as the CHECKER only,
        outside every timed region."""
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_py
            ekf, N = self.ekf, self.N
            scans = self.scans(m)[:nscan]
            rng = np.random.Generator(np.random.MT19937(99))
            lms = sorted(set(int(j) for _, ids in scans for j in ids) | set(int(j) for j in rng.choice(N, 48, replace=False) + 1))
            idx = [0, 1, 2] + [c for j in lms for c in (3 + 2 * (j - 1), 4 + 2 * (j - 1))]
            slot = {j: k + 1 for k, j in enumerate(lms)}  # global landmark id -> oracle map slot
            X = ekf.X
            orc = oracle_py.OracleEKF(self.ctx.cs.FLAG_INTENDED)
            orc.reset(X[idx], ekf.cov_gather(idx))
            same_idx = True
            for f in (orc, ekf):  # two control steps first: predict + heading (EKF.cpp:406-455, 328-352)
                for k in range(2):
                    f.predict(20.0, 0.01 * (k + 1), QE_BENCH, 73.0, 0.01)
                    f.observeHeading(1e-4, True)
            for Z, ids in scans:
                jo = orc.gate(Z, RE, GATE1, GATE2)[0]
                jo_glob = np.array([lms[j - 1] if j > 0 else 0 for j in jo], dtype=np.int32)
                if batch:
                    jg = ekf.gate(Z, RE, GATE1, GATE2)[0]
                    ekf.update(Z[:, jg > 0], RE, jg[jg > 0], True)
                    orc.update(Z[:, jo > 0], RE, jo[jo > 0], True)
                else:
                    jg = ekf.scan(Z, RE, GATE1, GATE2)[0]
                    orc.update(Z[:, jo > 0], RE, jo[jo > 0], False)
                same_idx = same_idx and bool(np.array_equal(jg, jo_glob)) and bool(np.array_equal(jg, ids))
            Xg, Pg = ekf.X[idx], ekf.cov_gather(idx)
            Xo, Po = orc.X, orc.P
            iu = np.triu_indices(len(idx))
            ex = float(np.max(np.abs(Xg - Xo)) / max(np.max(np.abs(Xo)), 1e-300))
            ep = float(np.max(np.abs(Pg[iu] - Po[iu])) / max(np.max(np.abs(Po[iu])), 1e-300))
            return {"checked_against": "CPU oracle (oracle/slam_oracle.hpp, FP64) on the marginal of the pose, the observed "
                                       f"landmarks and 48 random landmarks ({len(idx)} state entries, full covariance block)",
                    "calls": f"2 x (predict + observeHeading), {nscan} scans of {m} observations "
                             f"({'gate + joint batchUpdate' if batch else 'fused gate + sequential update'})",
                    "state_max_rel_err": ex, "cov_max_rel_err": ep, "indices_equal": same_idx,
                    "tolerance": 1e-9, "ok": bool(ex < 1e-9 and ep < 1e-9 and same_idx)}
        except Exception as e:  # pragma: no cover - reported, never fatal for the timing
            return {"ok": False, "error": f"{type(e).__name__}: {e}"}

    # ---- the per-scan chain (gate -> column snapshot -> gains) on its own stream, timed scan by scan ----
    def chain_probe(self, m, reps=8):
        """CUDA events on the chain stream around single scans.  From a flushed state the scans that fill the first
        bank run with no covariance pass in flight (the last one launches the pass on the pass stream, which the
        chain does not wait for); the scan after it runs while that pass works on the covariance next to it.  The
        pass hides behind the chain or the chain behind the pass: per scan the filter costs max(chain, pass share)."""
        ctx, ekf, torch = self.ctx, self.ekf, self.ctx.torch
        if m != 4:
            return None
        scans = self.scans(m)
        res = {"idle": [], "beside_pass": []}
        per_bank = None
        with torch.cuda.stream(self.stream):
            for r in range(reps):
                ekf.flush()
                ekf.sync()
                ctx.barrier()
                p0 = ekf.pass_count()[0]
                ev = [torch.cuda.Event(enable_timing=True)]
                ev[0].record(self.stream)
                launched_at = None
                for k in range(20):
                    ekf_scan(ekf, scans[(5 * r + k) % len(scans)][0], False, want_indices=False)
                    ev.append(torch.cuda.Event(enable_timing=True))
                    ev[-1].record(self.stream)
                    if launched_at is None and ekf.pass_count()[0] > p0:
                        launched_at = k
                    elif launched_at is not None:
                        break
                ekf.flush()
                ekf.sync()
                if launched_at is None:
                    continue
                per_bank = launched_at + 1
                if r > 0:
                    res["idle"].append(ev[0].elapsed_time(ev[1]))
                    res["beside_pass"].append(ev[launched_at + 1].elapsed_time(ev[launched_at + 2]))
        if not res["idle"]:
            return None
        med = {k: float(np.median(v)) for k, v in res.items()}
        med = dict(zip(med.keys(), ctx.allmax(*med.values())))
        lo = {k: float(np.min(v)) for k, v in res.items()}
        lo = dict(zip(lo.keys(), ctx.allmax(*lo.values())))
        return {"ms_per_scan_chain_idle": med["idle"], "ms_per_scan_chain_beside_pass": med["beside_pass"],
                "min_idle": lo["idle"], "min_beside_pass": lo["beside_pass"], "reps": len(res["idle"]),
                "scans_per_bank": per_bank,
                "note": "gate + column snapshot (peer push when sharded) + gains of one 4-observation scan, max over ranks; "
                        "idle: first scan after a flush; beside_pass: first scan after a full bank launched its pass "
                        "(it also carries the most pending terms)"}

    # ---- device-timed throughput (state resident in HBM) + live kernel timing ----
    def timed(self, m, steps, warmup, batch=False, strict=False):
        ctx, ekf, torch = self.ctx, self.ekf, self.ctx.torch
        scans = self.scans(m)
        for s in range(warmup):
            ekf_scan(ekf, scans[s % len(scans)][0], batch)
            if strict:
                ekf.flush()
        ekf.sync()
        sampler = ClockSampler(ctx.local)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.barrier()
        launches0 = ctx.lib.cslam_kernel_launches()
        passes0 = ekf.pass_count()[0]
        ekf.profile_begin(steps * max(1, m) + 8)
        sampler.start()
        updates = 0
        assoc0 = 0 if batch else ekf.scan_associations()
        with torch.cuda.stream(self.stream):
            ev0.record(self.stream)
            for s in range(steps):
                ekf_scan(ekf, scans[(warmup + s) % len(scans)][0], batch, want_indices=False)
                if strict:
                    ekf.flush()
                updates += (1 if batch else 0)
            ekf.flush()  # every deferred covariance term is applied INSIDE the timed region
            ev1.record(self.stream)
        ctx.barrier()
        if not batch:  # updates applied = observations the gate associated, counted on the device
            updates = ekf.scan_associations() - assoc0
        clocks = sampler.stop()
        cov_ms, cov_launches, cov_bytes = ekf.profile_end()
        launches = ctx.lib.cslam_kernel_launches() - launches0
        ms = ev0.elapsed_time(ev1)
        skipped = ekf.sync()
        ms, = ctx.allmax(ms)
        if ctx.world > 1:
            u, l = ctx.allsum(updates, launches)
            launches = int(l)
            if not ctx.sharded:
                updates = int(u)
        return {"ms": ms, "updates": updates, "cov_ms": cov_ms, "cov_launches": cov_launches, "cov_bytes": cov_bytes,
                "launches": int(launches), "clocks": clocks, "skipped": skipped, "steps": steps,
                "passes": ekf.pass_count()[0] - passes0}

    # ---- end to end through the public API: host observations in, state + indices out ----
    def e2e(self, m, steps, batch=False):
        ctx, ekf, torch = self.ctx, self.ekf, self.ctx.torch
        scans = self.scans(m)
        X_host = None
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        upd = 0
        with torch.cuda.stream(self.stream):
            e0.record(self.stream)
            for s in range(steps):
                Z = scans[(s + 1) % len(scans)][0]
                jb = ekf_scan(ekf, Z, batch)
                X_host = ekf.X  # D2H read of the step's result (n doubles through pinned staging)
                upd += (1 if batch else int((jb > 0).sum()))
            ekf.flush()
            e1.record(self.stream)
        ctx.barrier()
        ms = e0.elapsed_time(e1)
        ms, = ctx.allmax(ms)
        if ctx.world > 1 and not ctx.sharded:
            upd = int(ctx.allsum(upd)[0])
        return {"ms": ms, "updates": upd, "X": X_host}

    # ---- a whole drive cycle of test/main.cpp: 6 control steps (predict + observeHeading) + one scan ----
    def drive(self, m, cycles=4):
        ctx, ekf, torch = self.ctx, self.ekf, self.ctx.torch
        scans = self.scans(m)

        def cycle(Z):
            ekf.controlSteps(np.zeros(6), np.zeros(6), np.zeros(6), True, QE_BENCH, 73.0, 0.01, want_trace=False)
            ekf_scan(ekf, Z, want_indices=False)
        cycle(scans[0][0])
        ekf.sync()
        p0 = ekf.pass_count()[0]
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(self.stream):
            d0.record(self.stream)
            for c in range(cycles):
                cycle(scans[(c + 2) % len(scans)][0])
            ekf.flush()
            d1.record(self.stream)
        torch.cuda.synchronize()
        return {"control_steps_per_cycle": 6, "observations_per_cycle": m, "cycles_timed": cycles,
                "ms_per_cycle": d0.elapsed_time(d1) / cycles,
                "covariance_passes_per_cycle": (ekf.pass_count()[0] - p0) / cycles,
                "round1_ms_per_cycle": {"stepwise (7 passes)": 14.4, "merged heading passes (2 passes)": 4.6},
                "note": "heading and landmark updates are deferred rank-1 terms: one covariance pass per bank of up to 64 of them "
                        "(14 per cycle), the flush at the end of the timed cycles included"}


def ekf_roofline(ctx, n, N, m, t, batch, strict, peak, peak_src):
    shards = ctx.world if ctx.sharded else 1
    cov_ms, cov_launches, cov_bytes = t["cov_ms"], t["cov_launches"], t["cov_bytes"]
    ach = (cov_bytes / cov_launches) / (cov_ms / cov_launches * 1e-3) / 1e9 if cov_launches else 0.0
    upd_local = t["updates"] if (ctx.sharded or ctx.world == 1) else t["updates"] / ctx.world
    if batch:
        r_rank = 2 * m
        flops = float(r_rank) * n * (n + 1) / shards
        tpeak, tsrc = dmma_peak(ctx.local)
        t_launch = cov_ms / max(1, cov_launches) * 1e-3
        return {
            "bound": "tensor", "kernel": "k_cov_update_dmma (slam.h:260, rank-2m update on FP64 tensor cores, "
                                         "DMMA.8x8x4) incl. its panel-tiling kernel",
            "achieved": flops / t_launch / 1e12 if cov_launches else 0.0, "peak": tpeak, "unit": "TFLOP/s",
            "frac": (flops / t_launch / 1e12) / tpeak if cov_launches else 0.0, "peak_source": tsrc,
            "traffic": ncu_traffic("k_cov_update_dmma", n) if ctx.world == 1 else None,
            "algorithmic_flops_per_launch": flops, "algorithmic_bytes_per_launch": 8.0 * n * (n + 1) / shards,
            "hbm_gbs_same_launch": (cov_bytes / cov_launches) / t_launch / 1e9 if cov_launches else 0.0,
            "hbm_frac_same_launch": ((cov_bytes / cov_launches) / t_launch / 1e9) / peak if cov_launches else 0.0,
            "per": "GPU", "launches_timed": cov_launches, "avg_launch_ms": cov_ms / max(1, cov_launches),
        }
    rows = 2 * (1 if strict else min(m, 8))
    per_pass = upd_local / max(1, cov_launches)
    alg_update_bytes = (8.0 * n * (n + 1)) / shards + 112.0 * n  # cov R+W + 5 P columns + X + gating
    lazy = n >= 2048 or ctx.world > 1
    if lazy and per_pass > 8.5:
        # banks of more than 16 panel rows go through the FP64 tensor-core kernel (ekf_dmma.cu, out of place): the
        # pass is bound by the DMMA pipe, not by HBM — report it against the in-run DMMA peak, HBM figures beside it
        r_rank = 2.0 * per_pass
        flops = r_rank * n * (n + 1) / shards
        tpeak, tsrc = dmma_peak(ctx.local)
        t_launch = cov_ms / max(1, cov_launches) * 1e-3
        return {
            "bound": "tensor",
            "kernel": "k_cov_update_dmma (slam.h:260 for every pending update of a bank — up to 64 panel rows = 32 sequential "
                      "landmark updates per launch — rank-r term on the FP64 tensor cores (DMMA.8x8x4), the covariance read "
                      "from one array and written to its twin once per bank) incl. its panel-tiling kernel",
            "updates_per_launch": per_pass, "panel_rows_per_update": 2,
            "achieved": flops / t_launch / 1e12, "peak": tpeak, "unit": "TFLOP/s",
            "frac": (flops / t_launch / 1e12) / tpeak, "peak_source": tsrc,
            "traffic": ncu_traffic("k_cov_update_dmma", n) if ctx.world == 1 else None,
            "algorithmic_flops_per_launch": flops, "algorithmic_bytes_per_launch": 8.0 * n * (n + 1) / shards,
            "hbm_gbs_same_launch": ach, "hbm_frac_same_launch": ach / peak if peak else 0.0,
            "per": "GPU", "launches_timed": cov_launches, "avg_launch_ms": cov_ms / max(1, cov_launches),
        }
    if lazy and per_pass <= 1.01:
        kname = "k_cov_update"  # one- and two-row passes: the FMA streaming kernel, in place (ekf_lazy.cuh: lazy_flush)
    else:
        kname = "k_cov_update_tma_dense" if lazy else "k_cov_update_multi"
    return {
        "bound": "hbm",
        "kernel": (f"{kname} (slam.h:260 for every pending update — up to 16 panel rows = 8 sequential landmark updates "
                   f"per launch — in ONE tensor-map TMA read + write of the upper triangle, rank-r term on the FP64 "
                   f"tensor cores)" if kname == "k_cov_update_tma_dense" else
                   f"{kname} (slam.h:260, upper-triangle update, FMA streaming kernel)"),
        "updates_per_launch": per_pass, "panel_rows_per_update": 2,
        "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak if peak else 0.0,
        "peak_source": peak_src,
        "traffic": ncu_traffic("k_cov_update<2,128>" if kname == "k_cov_update" else kname, n) if ctx.world == 1 else None,
        "algorithmic_bytes_per_launch": 8.0 * n * (n + 1) / shards, "per": "GPU",
        "launches_timed": cov_launches, "avg_launch_ms": cov_ms / max(1, cov_launches),
        "whole_update_frac": (alg_update_bytes * upd_local / (t["ms"] * 1e-3) / 1e9) / peak if peak else 0.0,
        "whole_update_frac_note": "SURVEY §8d bytes assume one covariance pass per update; deferred passes apply "
                                  f"{per_pass:.1f} updates per pass, hence > 1" if per_pass > 1.01 else None,
    }


def ekf_config(N, m, world, sharded, batch=False, strict=False):
    """The `config` object of an EKF line — shared by our arm and the reference arm (the driver compares them)."""
    n = 3 + 2 * N
    upd_kind = (f"batched JOINT update of {m} observations per scan (rank {2 * m}): gate + stacked gain + "
                f"tensor-core covariance update" if batch else
                f"sequential update: gate + gain + covariance, {m} observation{'s' if m > 1 else ''} per scan" +
                (", one covariance pass per update (flush after every scan)" if strict else ""))
    return {
        "workload": (f"EKF-SLAM {N} landmarks (state dim {n}, FP64 P {8.0 * n * n / 1e9:.2f} GB), "
                     f"range-bearing observations, {upd_kind}" +
                     (f", covariance row-sharded over {world} GPUs" if sharded else
                      (f", {world} independent filter replicas" if world > 1 else ""))),
        "landmarks": N, "state_dim": n, "obs_per_step": m, "mode": "INTENDED (SURVEY Appendix A)",
        "l2": f"inputs larger than L2 ({4.0 * n * n / 1e9 / (world if sharded else 1):.1f} GB of upper "
              f"triangle streamed per GPU per covariance pass)",
        "parallelism": ("row-sharded covariance (block-cyclic 128-row tiles), observed columns exchanged over NVLink "
                        "peer memory inside the snapshot kernel, gains / gating overlapped with the covariance pass"
                        if sharded else
                        ("replicas only (one independent filter per GPU)" if world > 1 else "single GPU")),
    }


def ekf_result(ctx, eb, m, steps, warmup, batch=False, strict=False, with_e2e=True, with_parity=True):
    """One EKF workload on an existing EkfBench: parity check, device-timed loop, end-to-end loop -> result dict."""
    N, n = eb.N, eb.n
    par = eb.parity(m, batch) if with_parity else None
    t = eb.timed(m, steps, warmup, batch, strict)
    e = eb.e2e(m, steps, batch) if with_e2e else None
    if ctx.rank != 0:
        return None
    peak, peak_src = measured_peaks()
    value = t["updates"] / (t["ms"] * 1e-3)
    sharded, world = ctx.sharded, ctx.world
    out = {
        "metric": "EKF updates/sec", "value": value, "unit": "updates/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": t["ms"] / steps, "higher_is_better": True,
        "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": ekf_config(N, m, world, sharded, batch, strict),
        "clocks": t["clocks"],
        "gpu_launches": t["launches"],
        "covariance_passes": t["passes"],
        "roofline": ekf_roofline(ctx, n, N, m, t, batch, strict, peak, peak_src),
        "skipped_updates": t["skipped"] & ((1 << 20) - 1),
        # the device status word counts 1 << 20 per wait of a sharded column exchange that gave up on a peer (the
        # numbers of such a run are void)
        "peer_exchange_timeout": bool(t["skipped"] >= (1 << 20) or t["skipped"] < 0),
    }
    if e is not None:
        out["e2e"] = {"value": e["updates"] / (e["ms"] * 1e-3), "unit": "updates/s",
                      "h2d_bytes_per_step": int(2 * m * 8 + 4 * 8 + 2 * 8 + m * 4 + 4 * 8),
                      "d2h_bytes_per_step": int(n * 8 + m * (4 + 8 + 8))}
        out["state_checksum"] = float(np.sum(e["X"][:3])) if e["X"] is not None else None
    if par is not None:
        out["parity_check"] = par
    if batch:
        out["observations_per_s"] = value * m
    return out


def c1_loop():
    """C1 of BASELINE.json: the reference's own simulated waypoint loop (test/main.cpp:132-200, 30-landmark world,
    22,018 control steps, 3,259 observation steps) replayed by the C++ host driver through the drop-in adaptor:
    stepwise calls vs the fused forms (k control steps per launch, cslam_ekf_control_steps; joint update +
    augmentation per launch, cslam_ekf_observe_step).  Beside it the reference's own binary (test/main.cpp
    compiled unmodified, oracle/_ref/slam_ref, stdout to /dev/null) on one host core — a baseline, not a target:
    at n = 53 everything is launch- and PCIe-latency-bound."""
    import re
    import subprocess
    exe = os.path.join(ROOT, "conan_slam_b200", "lib", "sim_main")
    if not os.path.exists(exe):
        return {"skipped": "conan_slam_b200/lib/sim_main not built"}
    out = {"control_steps": None}

    def run(flags):
        r = subprocess.run([exe, "--print-every", "0"] + flags, capture_output=True, text=True, timeout=300)
        m = re.search(r"done: (\d+) control steps, n=(\d+), skipped updates=(\d+), loop wall time ([0-9.]+) s", r.stdout)
        if r.returncode != 0 or not m:
            raise RuntimeError(f"sim_main {flags}: rc={r.returncode} {r.stdout[-200:]} {r.stderr[-200:]}")
        out["control_steps"], out["final_state_dim"] = int(m.group(1)), int(m.group(2))
        return float(m.group(4))
    out["loop_s_stepwise_calls"] = run([])
    out["loop_s_fused"] = run(["--fused"])
    out["loop_s_fused_no_cov_writeback"] = run(["--fused", "--no-cov-writeback"])
    out["control_steps_per_s_fused"] = out["control_steps"] / out["loop_s_fused"]
    ref = os.path.join(ROOT, "oracle", "_ref", "slam_ref")
    if os.path.exists(ref):
        t0 = time.perf_counter()
        with open(os.devnull, "w") as dn:
            rc = subprocess.run([ref], stdout=dn, stderr=dn, timeout=600).returncode
        out["cpu_baseline"] = {"value": time.perf_counter() - t0, "unit": "s per loop (whole process)", "cores": 1,
                               "kind": "reference", "rc": rc,
                               "sample": "the reference's own test/main.cpp + EKF.cpp (oracle/_ref/slam_ref, FP32, "
                                         "Eigen stand-in), full loop, stdout discarded"}
    return out


def guarded(name, fn):
    """Extras never take the headline down: failures are reported in place."""
    try:
        t0 = time.time()
        r = fn()
        if isinstance(r, dict):
            r["wall_s"] = round(time.time() - t0, 1)
        return r
    except Exception as e:  # pragma: no cover
        import traceback
        log(f"[bench] extra '{name}' failed: {traceback.format_exc()}")
        return {"error": f"{type(e).__name__}: {e}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--landmarks", type=int, default=20000)
    ap.add_argument("--obs", type=int, default=None, help="observations per scan (default 4; 32 with --batch)")
    ap.add_argument("--batch", action="store_true",
                    help="C3 of BASELINE.json: one JOINT update of all observations of a scan (rank 2m, FP64 tensor cores)")
    ap.add_argument("--strict", action="store_true", help="one covariance pass per scan (flush after every scan)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline only (no C2 / C3-batch / per-observation / PF / C5 blocks)")
    ap.add_argument("--multi", default="sharded", choices=["sharded", "replicas"],
                    help="N>1: row-sharded covariance of ONE filter (strong scaling) or independent filter replicas")
    ap.add_argument("--workload", default="ekf", choices=["ekf", "pf"],
                    help="ekf: the headline (BASELINE.json metric, first half); pf: particle-steps/sec (C4)")
    ap.add_argument("--particles", type=int, default=1 << 20)
    ap.add_argument("--pf-landmarks", type=int, default=500)
    ap.add_argument("--pf-obs", type=int, default=4)
    ap.add_argument("--pf-spread", default="balanced", choices=["balanced", "realistic", "adversarial"])
    args = ap.parse_args()
    _capture_stdout()
    if args.obs is None:
        args.obs = 32 if args.batch else 4
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return
    if args.workload == "pf":
        args.pf_obs = args.obs if args.obs != 32 else args.pf_obs
        run_pf(args)
        return

    ctx = Ctx(args)
    N, m = args.landmarks, args.obs
    extras_on = not args.no_extras and not args.batch and not args.strict and args.multi == "sharded" and N == 20000
    eb = EkfBench(ctx, N, max(8, args.steps + args.warmup))
    out = ekf_result(ctx, eb, m, args.steps, args.warmup, batch=args.batch, strict=args.strict)
    extras = {}
    if not args.batch and not args.strict:
        cp = guarded("chain_probe", lambda: eb.chain_probe(m))
        if ctx.rank == 0 and cp is not None:
            out["chain_probe"] = cp
    if ctx.world == 1 and not args.batch:
        d = guarded("drive_cycle", lambda: eb.drive(m))
        if ctx.rank == 0:
            out["drive_cycle"] = d
    if extras_on:
        # ---- the rest of the measurement contract in the same run ----
        if ctx.world == 1:
            extras["c3_obs1"] = guarded("c3_obs1", lambda: ekf_result(ctx, eb, 1, 12, 3, strict=True, with_parity=False))
            extras["c3_batch"] = guarded("c3_batch", lambda: ekf_result(ctx, eb, 32, 6, 3, batch=True))
            extras["c3_obs8"] = guarded("c3_obs8", lambda: ekf_result(ctx, eb, 8, 24, 4, with_parity=False, with_e2e=False))
    eb.close()
    if extras_on:
        if ctx.world == 1:
            def c2():
                e2 = EkfBench(ctx, 2000, 64)
                r = ekf_result(ctx, e2, 4, 200, 10)
                e2.close()
                return r
            extras["c2"] = guarded("c2", c2)
            if ctx.rank == 0:
                extras["c1_loop"] = guarded("c1_loop", c1_loop)
        if ctx.world >= 2:
            def c5():
                e5 = EkfBench(ctx, 60000, 40)
                r = ekf_result(ctx, e5, 4, 32, 4)
                e5.close()
                return r
            free_gb = ctx.torch.cuda.mem_get_info()[0] / 1e9
            need_gb = 2 * 8.0 * 120003 ** 2 / 1e9 / ctx.world * 1.15
            if free_gb > need_gb / 2 + 4:  # one array is the minimum; the ping-pong twin is taken when it fits
                extras["c5"] = guarded("c5", c5)
            else:
                extras["c5"] = {"skipped": f"60,000 landmarks need {need_gb / 2:.0f} GB per GPU, {free_gb:.0f} GB free"}
        pf_steps = 6
        extras["pf_c4"] = guarded("pf_c4", lambda: pf_case(args, pf_steps, 3, "balanced", with_cpu=not args.no_cpu_baseline, quiet=True))
        extras["pf_c4_realistic"] = guarded("pf_c4_realistic", lambda: pf_case(args, pf_steps, 3, "realistic", with_cpu=False, quiet=True))
        if ctx.world > 1:
            extras["pf_c4_adversarial"] = guarded("pf_c4_adversarial", lambda: pf_case(args, pf_steps, 3, "adversarial", with_cpu=False, quiet=True))
    if ctx.rank == 0:
        if extras:
            out["extras"] = {k: v for k, v in extras.items() if v is not None}
        if not args.no_cpu_baseline and ctx.world == 1:
            t1 = time.time()
            # the reference's own compiled sources (oracle/_ref) when they travelled with the snapshot,
            # else the oracle port; the multi-threaded port beside it as a generous baseline
            out["cpu_baseline"] = ref_ekf_update_rate(N, budget_s=12.0) or \
                cpu_ekf_update_rate(N, 1, slab_cols=3000, reps=2)
            mt = cpu_ekf_update_rate(N, os.cpu_count() or 1, slab_cols=3000, reps=2)
            out["cpu_baseline"]["oracle_port_all_cores"] = {
                "cores": mt["cores"], "value": mt["value"], "kind": "port",
                "update_only_updates_per_s": mt["update_only_updates_per_s"]}
            out["cpu_baseline"]["extrapolated"] = True
            log(f"[bench] cpu baseline took {time.time() - t1:.1f}s")
        emit(out)
    if ctx.world > 1:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
