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
    explicit ProfScope(cslam_ekf* h_) : h(h_), on(h_->prof && h_->prof_used + 2 <= (int)h_->prof_ev.size()) {
        if (on) cudaEventRecord(h->prof_ev[h->prof_used], h->stream);
    }
    ~ProfScope() {
        if (on) {
            cudaEventRecord(h->prof_ev[h->prof_used + 1], h->stream);
            h->prof_used += 2;
            h->prof_bytes += 8.0 * (double)h->n * ((double)h->n + 1.0) / (double)h->sh.world;
        }
    }
};

}  // namespace cslam
