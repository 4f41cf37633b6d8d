// ekf_handle.cuh — the EKF handle behind cslam_ekf_t, shared by ekf.cu and ekf_sharded.cu.
#pragma once
#include <vector>

#include "common.cuh"
#include "nccl_dl.cuh"

namespace cslam {

// ------------------------------------------------------------------------------------
// Handle
// ------------------------------------------------------------------------------------
constexpr int kMaxRank = 2 * CSLAM_MAX_BATCH_OBS;  // 64

struct BatchSmall {  // device-resident scratch of the joint update (EKF.cpp:93-129)
    double hu[CSLAM_MAX_BATCH_OBS][2][3];
    double lu[CSLAM_MAX_BATCH_OBS][2][2];
    double V[kMaxRank];
    double G[kMaxRank * kMaxRank];  // row-major r x r: L^-1 (literal) or L^-T (Q1 intended)
    double u[kMaxRank];             // G * G^T * V
    int f[CSLAM_MAX_BATCH_OBS];
};

// Deferred covariance passes ("lazy" engine, large and sharded maps; ekf_lazy.cuh).  The big array holds
//   P_dev = P_true + (sum of the pending rank-1 terms)  for rows >= 3,
// while X, rows 0..2 (R3) and the landmarks' 2x2 diagonal blocks (D) are always current.  Heading updates and
// landmark updates append their panel rows to a BANK of kLazyBank rows; one TMA / tensor-core pass
// (cov_tma.cu) applies a whole bank — e.g. the six heading updates and the four observations of one
// test/main.cpp drive cycle in ONE read + write of the covariance instead of two (round 1) or ten (reference).
// Gains read the few entries of P they need from a column snapshot and bring them up to date with the
// pending terms themselves.  The pass runs on its own stream and overlaps the gate / gain chain of the
// following scans; with a ping-pong pair of arrays the chain never waits for a pass that is still running.
// Panel rows per bank = rank of one pass.  Up to 16 rows go through the TMA streaming pass (cov_tma.cu, HBM-bound,
// its time does not depend on the rank); a bank of up to 64 rows through the FP64 tensor-core kernel of the joint
// update (ekf_dmma.cu, out of place into the ping-pong twin): 3.4 ms for 32 sequential updates at N = 20 000 against
// 2.1 ms for 8 — the covariance is read and written once per 32 updates instead of once per 8.  LazyState::bank_rows
// is the size in use (CSLAM_LAZY_BANK overrides); kLazyBankMax sizes the buffers.
constexpr int kLazyBankMax = 64;
constexpr int kLazyBankTma = 16;
constexpr int kSeqGroupLazyMax = 8;  // observations per snapshot group (2 columns each)
struct GroupHeader {  // written by the snapshot kernel, read by k_gain_group_lazy (ekf_lazy.cuh)
    int f[kSeqGroupLazyMax];                              // first state index of observation k, -1 = none
    double Ac[2 * kLazyBankMax][3 + 2 * kSeqGroupLazyMax];   // pending term t at column b of the group's marginal
};
struct LazyState {
    bool on = false;
    bool pingpong = false;
    double* Pbuf[2] = {nullptr, nullptr};  // Pbuf[0] == h->P; Pbuf[1] only with pingpong
    unsigned char map[2][128];             // tensor maps of Pbuf[0 / 1] (64 bytes used, copied by value per launch)
    int stable = 0;   // array the chain stream may read: never written by a pass in flight when pingpong
    int newest = 0;   // array that holds the newest state once every launched pass has completed
    int bank_rows = kLazyBankTma;  // rows per bank in use (<= kLazyBankMax)
    int bank = 0;     // bank that receives new panel rows: rows [bank * kLazyBankMax, +bank_rows) of h->A
    int np = 0;       // rows pending in `bank`
    unsigned long long eps_mask = 0;  // bit k: pending row k is a heading term (diagonal += FLT_MIN, slam.h:719)
    int infl_rows = 0;           // rows of the OTHER bank that the pass in flight applies and `stable` lacks
    unsigned long long infl_eps_mask = 0;
    double* pass_panels = nullptr;    // pre-tiled panel copy for the tensor-core pass (ekf_dmma.cu), pass stream only
    bool pass_pending_wait = false;  // a pass was launched and the chain stream has not waited for it yet
    bool stable_busy = false;        // ... and that pass works IN PLACE on `stable` (no ping-pong twin, or a small pass)
    cudaStream_t pass_stream = nullptr;
    cudaEvent_t ev_chain = nullptr;  // "every reader of the array the next pass overwrites is done"
    cudaEvent_t ev_pass = nullptr;   // completion of the most recently launched pass
    int num_sms = 0;
    int stages = 2;  // ring depth of the TMA pass (direct-store mode: 2 x 32 KB measured best on a B200)
    unsigned long long passes = 0;   // passes launched since create (diagnostics)
    // Sharded handles, column snapshot over NVLink peer memory (no collective on the chain): every rank WRITES the
    // entries it stores of the observed columns straight into the snapshot buffer of every rank (CUDA-IPC
    // mapped), then raises a flag on every peer; a one-warp kernel waits for all flags before the gains read.
    bool peers_ready = false;
    double* xbuf = nullptr;           // [2][2 * kSeqGroupLazyMax][lda]: snapshot buffers (double-buffered by epoch parity)
    size_t xll_off = 0;               // byte offset of the flagged-cell buffers ([2][2 * kSeqGroupLazyMax][lda] x 16 B) in xbuf's allocation
    unsigned long long* sig = nullptr;  // [8]: sig[q] = last epoch rank q has finished pushing to this rank
    double* peer_xbuf[8] = {};        // IPC mappings (self = own pointers)
    unsigned long long* peer_sig[8] = {};
    void** d_peer_tab = nullptr;      // device copy: [0..7] xbuf pointers, [8..15] sig pointers
    unsigned long long epoch = 0;     // snapshots taken so far (same on every rank: SPMD)
    unsigned* push_ticket = nullptr;  // device: last-block ticket of the push kernel
    // twins of R3 / D: the fused group-gain kernel (k_gain_group_lazy) writes the followed panels there while other
    // blocks still read the old ones; the handle's R3 / D pointers are swapped after the launch
    double* R3alt = nullptr;
    double* Dalt = nullptr;
    GroupHeader* hdr = nullptr;
    // CSLAM_KTRACE=<file>: device timestamps (globaltimer) of the snapshot and group-gain kernels, 4 per group:
    // [snapshot first block in, snapshot last block out, group first block in, group last block out]; dumped at destroy
    unsigned long long* ktrace = nullptr;
    int kslot = 0;
    static constexpr int kTraceCap = 8192;
    bool defer_gate_merge = false;    // CSLAM_GATE_DEFER=1: the snapshot kernel merges the gate candidates (measured slower beside a pass: 0.128 vs 0.108 ms per scan)
    bool fused_gains = true;          // CSLAM_GAIN_FUSED=0: one gain kernel per observation + follow (regression tests)
};

struct GateScratch {
    double* part_nd = nullptr;   // [blocks][m]
    double* part_out = nullptr;  // [blocks][m]
    int* part_j = nullptr;       // [blocks][m]
    int* d_jbest = nullptr;      // [CSLAM_MAX_OBS]
    double* d_nbest = nullptr;
    double* d_outer = nullptr;
    int max_blocks = 0;
};

}  // namespace cslam

struct cslam_ekf {
    int device = 0;
    unsigned flags = 0;
    int cap_landmarks = 0;
    int n_cap = 0;
    size_t ld = 0;   // leading dimension of P (doubles)
    size_t lda = 0;  // leading dimension of A / PHT panels (doubles)
    int n = 3;
    int cur = 0;  // which X buffer is current
    double* X[2] = {nullptr, nullptr};
    double* P = nullptr;
    double* A = nullptr;    // [kMaxRank][lda]
    double* PHT = nullptr;  // [kMaxRank][lda]
    double* dmma_panels = nullptr;  // pre-tiled panel copies of the tensor-core joint update (lazy)
    cslam::BatchSmall* small = nullptr;
    int* status = nullptr;        // device: #skipped updates
    unsigned* ticket = nullptr;   // device: last-block tickets (predict, gate)
    cslam::GateScratch gate;
    unsigned long long* assoc_count = nullptr;  // device: observations associated by fused scans since create
    void* pinned = nullptr;  // host staging
    size_t pinned_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t scan_ev = nullptr;  // marks "association indices of the last fused scan are in pinned memory"
    bool own_stream = false;
    // row sharding over GPUs (world == 1: single GPU, R3 aliases P, no NCCL)
    cslam::Shard sh = {0, 1};
    ncclComm_t comm = nullptr;
    int local_rows_cap = 0;      // rows of P stored on this rank
    double* R3 = nullptr;        // rows 0..2 of P, replicated on every rank ([3][ld]; rank 0: alias of P)
    double* colbuf = nullptr;    // [kMaxRank][lda] column exchange buffer (all-reduced)
    double* D = nullptr;         // [3][dcap] replicated cache of the 2x2 diagonal blocks (gating)
    int dcap = 0;
    bool diag_dirty = true;
    cslam::LazyState lz;
    double* trace_dev = nullptr;  // pose trace of cslam_ekf_control_steps (grown on demand, never per call)
    size_t trace_cap = 0;
    double* acc_dev = nullptr;    // scratch of the accessors (cov block / landmark marginals), grown on demand
    size_t acc_cap = 0;
    bool attr_chol = false, attr_rank = false, attr_step = false;  // cudaFuncSetAttribute done for this handle's device
    // diagnostics: event pairs around covariance-update launches
    bool prof = false;
    std::vector<cudaEvent_t> prof_ev;
    int prof_used = 0;
    double prof_bytes = 0.0;
};

namespace cslam {

struct ProfScope {  // records start/stop events around one covariance-update launch when profiling
    cslam_ekf* h;
    bool on;
    cudaStream_t st;
    explicit ProfScope(cslam_ekf* h_, cudaStream_t stream = nullptr)
        : h(h_), on(h_->prof && h_->prof_used + 2 <= (int)h_->prof_ev.size()), st(stream ? stream : h_->stream) {
        if (on) cudaEventRecord(h->prof_ev[h->prof_used], st);
    }
    ~ProfScope() {
        if (on) {
            cudaEventRecord(h->prof_ev[h->prof_used + 1], st);
            h->prof_used += 2;
            h->prof_bytes += 8.0 * (double)h->n * ((double)h->n + 1.0) / (double)h->sh.world;
        }
    }
};

}  // namespace cslam
