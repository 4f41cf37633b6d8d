// ekf.cu — EKF-SLAM hot path for sm_100a: predict, heading update, Kalman gain (sparse H),
// in-place symmetric covariance update over the upper triangle, in-place augmentation, joint
// (batch) update, plus the cslam_ekf_* C ABI (include/cslam.h).  Gating lives in gate.cu, the
// streaming covariance kernel in cov_update.cuh, the FP64 tensor-core kernel in ekf_dmma.cu.
//
// Data layout in HBM:
//   X      : 2 x n_cap doubles (ping-pong so gain/heading kernels never read what they write)
//   P      : rows x ld doubles, row-major, ld = n_cap+1 rounded up to 16 doubles (128 B rows);
//            ONLY the upper triangle (j >= i) is authoritative — the update streams
//            8*n*(n+1) bytes instead of 16*n^2.
//   A (W1) : r x lda doubles, one contiguous n-vector per update rank, r <= 64 (slam.h:257).
// Multi-GPU (world > 1): P is row-sharded in block-cyclic 128-row tiles (common.cuh Shard); rows
// 0..2 (R3) and X are replicated and updated redundantly with the same IEEE operations on every
// rank; per update only the observed landmark's two columns are exchanged (one all-reduce of
// 2*n doubles), then every rank forms the gain and streams its own rows.
#include <limits>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "cov_update.cuh"
#include "ekf_handle.cuh"

namespace cslam {

// ------------------------------------------------------------------------------------
// Kernels
// ------------------------------------------------------------------------------------

// Pvv <- Gv Pvv Gv^T + Gu Q Gu^T (EKF.cpp:439) and the pose advance (EKF.cpp:446-450), one thread.
__device__ void predict_pvv_pose(double* __restrict__ X, double* __restrict__ P, size_t ld, double v, double swa,
                                 double q00, double q01, double q10, double q11, double wb, double dt, double phi,
                                 double s, double c, double g02, double g12) {
    double Pv[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Pv[i][j] = psym(P, ld, i, j);
    const double Gv[3][3] = {{1, 0, g02}, {0, 1, g12}, {0, 0, 1}};
    const double Gu[3][2] = {{dt * c, -v * dt * s}, {dt * s, v * dt * c}, {dt * sin(swa) / wb, v * dt * cos(swa) / wb}};
    const double Q[2][2] = {{q00, q01}, {q10, q11}};
    double GP[3][3], GQ[3][2];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double a = 0;
            for (int k = 0; k < 3; k++) a += Gv[i][k] * Pv[k][j];
            GP[i][j] = a;
        }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 2; j++) {
            double a = 0;
            for (int k = 0; k < 2; k++) a += Gu[i][k] * Q[k][j];
            GQ[i][j] = a;
        }
    for (int i = 0; i < 3; i++)
        for (int j = i; j < 3; j++) {
            double a = 0, b = 0;
            for (int k = 0; k < 3; k++) a += GP[i][k] * Gv[j][k];
            for (int k = 0; k < 2; k++) b += GQ[i][k] * Gu[j][k];
            P[(size_t)i * ld + j] = a + b;
        }
    X[0] = X[0] + v * dt * c;
    X[1] = X[1] + v * dt * s;
    X[2] = pi2pi(phi + v * dt * sin(swa) / wb);
}

// EKF.cpp:406-455 predict on rows 0..2 (`P` here = the row panel R3: P itself on one GPU, the
// replicated panel when sharded).  Rows 0..1 for columns [3, 3+width) get Gv * (.) — with
// Gv = [[1,0,a],[0,1,b],[0,0,1]] that is row0 += a*row2, row1 += b*row2.  The mirrored columns
// (EKF.cpp:443) live in the lower triangle and are not stored.  The block that draws the last
// ticket updates Pvv and the pose, after every block has read the old phi.
__global__ void __launch_bounds__(256) k_predict(double* __restrict__ X, double* __restrict__ P, size_t ld, int n,
                                                 double v, double swa, double q00, double q01, double q10,
                                                 double q11, double wb, double dt, int width,
                                                 unsigned* __restrict__ ticket) {
    const double phi = X[2];
    const double s = sin(swa + phi), c = cos(swa + phi);
    const double g02 = -v * dt * s, g12 = v * dt * c;
    const int col = 3 + blockIdx.x * blockDim.x + threadIdx.x;
    if (col < 3 + width) {
        const double p0 = P[col], p1 = P[ld + col], p2 = P[2 * ld + col];
        P[col] = p0 + g02 * p2;
        P[ld + col] = p1 + g12 * p2;
    }
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last || threadIdx.x != 0) return;
    *ticket = 0;
    predict_pvv_pose(X, P, ld, v, swa, q00, q01, q10, q11, wb, dt, phi, s, c, g02, g12);
}

// EKF.cpp:328-352 + slam.h:700-725 with H = e_2^T.  Column 2 of P is (P[0][2], P[1][2],
// row 2 from the diagonal on) — all inside rows 0..2.  Writes Xout = Xin + W*v and the rank-1
// panel a = p/sqrt(S): for symmetric P the Joseph form C P C^T + W R W^T equals P - p p^T / S.
__global__ void __launch_bounds__(256) k_heading_gain(const double* __restrict__ Xin, double* __restrict__ Xout,
                                                      const double* __restrict__ P, size_t ld, int n,
                                                      double phi_meas, double R, double* __restrict__ A) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = pi2pi(phi_meas - Xin[2]);
    const double S = P[2 * ld + 2] + R;
    const double SI = 1.0 / S;
    const double p = i < 2 ? P[(size_t)i * ld + 2] : P[2 * ld + i];
    Xout[i] = Xin[i] + (p * SI) * v;
    A[i] = p / sqrt(S);
}

// Small maps (test/main.cpp's own 30-landmark world, n <= kSmallN): k consecutive control steps —
// predict (EKF.cpp:406-455) then observeHeading (EKF.cpp:328-352 -> slam.h:700-725) per step, the body
// of test/main.cpp:140-168 — in ONE launch of ONE CTA.  At n = 63 every per-step kernel is pure launch
// latency (first "next" row); here the phases of a step are separated by __syncthreads and
// the 32 KB covariance stays in L1/L2.  Same operations per element as k_predict / k_heading_gain /
// k_cov_update<1>, hence bit-identical results.
constexpr int kSmallN = 1024;
constexpr int kMaxControlSteps = 16;
struct ControlPack {
    double v[kMaxControlSteps], swa[kMaxControlSteps], phi[kMaxControlSteps];
    int k, use_heading, width;
    double q00, q01, q10, q11, wb, dt, r_heading;
};
// X and P are deliberately NOT __restrict__: threads of the CTA communicate through them across
// __syncthreads (with __restrict__ the compiler forwards a thread's own earlier load of X[2] past the barrier).
__global__ void __launch_bounds__(1024) k_control_steps(double* X, double* P, size_t ld, int n, ControlPack cp,
                                                        double* __restrict__ pose_trace) {
    __shared__ double sA[kSmallN];
    const int tid = threadIdx.x;
    for (int st = 0; st < cp.k; st++) {
        // ---- predict
        const double v = cp.v[st], swa = cp.swa[st];
        const double phi = X[2];
        const double s = sin(swa + phi), c = cos(swa + phi);
        const double g02 = -v * cp.dt * s, g12 = v * cp.dt * c;
        for (int col = 3 + tid; col < 3 + cp.width; col += blockDim.x) {
            const double p0 = P[col], p1 = P[ld + col], p2 = P[2 * ld + col];
            P[col] = p0 + g02 * p2;
            P[ld + col] = p1 + g12 * p2;
        }
        __syncthreads();  // every thread has read the old phi
        if (tid == 0) predict_pvv_pose(X, P, ld, v, swa, cp.q00, cp.q01, cp.q10, cp.q11, cp.wb, cp.dt, phi, s, c, g02, g12);
        __syncthreads();
        // ---- observeHeading: gain (k_heading_gain) ...
        if (cp.use_heading) {
            const double vinn = pi2pi(cp.phi[st] - X[2]);
            const double S = P[2 * ld + 2] + cp.r_heading;
            const double SI = 1.0 / S;
            double p = 0.0, x = 0.0;
            if (tid < n) {  // n <= kSmallN = blockDim.x
                p = tid < 2 ? P[(size_t)tid * ld + 2] : P[2 * ld + tid];
                x = X[tid];
            }
            __syncthreads();  // X[2] and column 2 are read before anyone overwrites them
            if (tid < n) {
                X[tid] = x + (p * SI) * vinn;
                sA[tid] = p / sqrt(S);
            }
            __syncthreads();
            // ... and the rank-1 pass over the upper triangle (k_cov_update<1>, diag_eps = FLT_MIN, slam.h:719)
            for (int idx = tid; idx < n * n; idx += blockDim.x) {
                const int i = idx / n, j = idx % n;
                if (j < i) continue;
                double o = P[(size_t)i * ld + j] - sA[i] * sA[j];
                if (j == i) o += kFltMin;
                P[(size_t)i * ld + j] = o;
            }
            __syncthreads();
        }
        if (pose_trace != nullptr && tid < 3) pose_trace[3 * st + tid] = X[tid];
        __syncthreads();
    }
}

// Per-observation prologue of slam.h:235-266 using the sparse H (robot block + one landmark
// block): only the 5x5 sub-block Pc of P at columns {0,1,2,f,f+1} enters S.
struct GainSmall {
    double H[2][5];
    double G[2][2];
    double V[2];
    int ok;
};
__device__ void gain_prologue(const double* X, const double (&Pc)[5][5], int f, double zr, double zb,
                              const double R[4], unsigned flags, GainSmall& g) {
    const ObsLin o = observe_lin(X[0], X[1], X[2], X[f], X[f + 1]);
    for (int a = 0; a < 2; a++) {
        for (int c = 0; c < 3; c++) g.H[a][c] = o.hu[a][c];
        g.H[a][3] = o.lu[a][0];
        g.H[a][4] = o.lu[a][1];
    }
    g.V[0] = zr - o.zr;
    g.V[1] = pi2pi(zb - o.zb);
    double PHTc[5][2];
    for (int a = 0; a < 5; a++)
        for (int k = 0; k < 2; k++) {
            double s = 0;
            for (int b = 0; b < 5; b++) s += Pc[a][b] * g.H[k][b];
            PHTc[a][k] = s;
        }
    double S[2][2];
    for (int k = 0; k < 2; k++)
        for (int l = 0; l < 2; l++) {
            double s = 0;
            for (int a = 0; a < 5; a++) s += g.H[k][a] * PHTc[a][l];
            S[k][l] = s + R[k + 2 * l];
        }
    // makeSymmetric (slam.h:247), LLT (slam.h:417-423), SCHOL.inverse() through PartialPivLU
    // (slam.h:251) — op for op as oracle::cholesky_update; failure -> zero gain (slam.h:252-255).
    double Ss[2][2], L[2][2], Li[2][2];
    for (int k = 0; k < 2; k++)
        for (int l = 0; l < 2; l++) Ss[k][l] = (S[k][l] + S[l][k]) * 0.5;
    int ok = chol_lower<2>(Ss, L) ? 1 : 0;
    inv_lu<2>(L, Li);
    for (int k = 0; k < 2; k++)
        for (int l = 0; l < 2; l++) ok = ok && isfinite(Li[k][l]);
    double G00 = 0, G01 = 0, G10 = 0, G11 = 0;
    if (ok) {
        if (flags & CSLAM_FLAG_Q1_METRIC_S) {  // G = L^-T
            G00 = Li[0][0]; G01 = Li[1][0]; G10 = Li[0][1]; G11 = Li[1][1];
        } else {  // literal: G = L^-1
            G00 = Li[0][0]; G01 = Li[0][1]; G10 = Li[1][0]; G11 = Li[1][1];
        }
    }
    g.G[0][0] = G00; g.G[0][1] = G01; g.G[1][0] = G10; g.G[1][1] = G11;
    g.ok = ok;
}

// slam.h:243,257-259 for one observation: PHT = P H^T (5 columns of P), W1 = PHT G,
// W = W1 G^T, Xout = Xin + W V; the rank-2 panel W1 (rows 2*kprev, 2*kprev+1 of A) feeds the
// covariance kernel.
// SH = false: columns f, f+1 of P are read in place (row part coalesced, column part strided).
// SH = true : they come from the all-reduced exchange buffer `colbuf`; rows 0..2 from `R3`.
// kprev > 0: the covariance passes of the kprev previous observations of this scan have NOT run yet
// (k_cov_update_multi applies them all in one pass afterwards); every entry of P read here is brought
// to its "after those updates" value by subtracting the pending rank-2 terms in update order — the
// same operations the covariance kernel will perform, so the values are bit-identical.
// idf_dev != nullptr: the association index is read from device memory (fused scan); 0 = skipped.
// The thread block computes the observation's small prologue (25 threads bring the 5x5 block of P up
// to date, one thread factorises), then every thread handles the state entries i = i0, i0 + stride, ...
// (A single-CTA variant looping over all observations of a scan was measured and dropped: at n = 4 003
// one CTA cannot keep enough loads in flight — 124 us per 4-observation scan against 91 us for four
// 16-CTA launches.)
template <bool SH>
__device__ __forceinline__ void gain_body(const double* Xin, double* Xout, const double* __restrict__ P,
                                          const double* __restrict__ R3, const double* __restrict__ colbuf,
                                          size_t ld, int n, double zr, double zb, int idf, const double (&R)[4],
                                          unsigned flags, double* A, size_t lda, int* __restrict__ status, int kprev,
                                          int i0, int stride, bool count_status, GainSmall& g, double (&sPc)[5][5]) {
    double* Aout = A + (size_t)2 * kprev * lda;
    if (idf == 0) {  // no landmark passed the gate: X is carried over, a zero panel is a no-op update
        for (int i = i0; i < n; i += stride) {
            Xout[i] = Xin[i];
            Aout[i] = 0.0;
            Aout[lda + i] = 0.0;
        }
        return;
    }
    const int f = 3 + 2 * (idf - 1);
    // stored value of P(r, c), r <= c, c in {0, 1, 2, f, f+1}
    auto praw = [&](int r, int c) -> double {
        if (r < 3) return R3[(size_t)r * ld + c];
        if constexpr (SH) {
            return colbuf[(size_t)(c - f) * lda + r];
        } else {
            return P[(size_t)r * ld + c];
        }
    };
    // P(r, c) after the kprev pending updates
    auto pfix = [&](double v, int r, int c) -> double {
        for (int q = 0; q < kprev; q++) {
            const double* a0 = A + (size_t)2 * q * lda;
            const double* a1 = a0 + lda;
            v = v - rank2_term(a0[r], a1[r], a0[c], a1[c]);
        }
        return v;
    };
    if (threadIdx.x < 25) {
        const int a = threadIdx.x / 5, b = threadIdx.x % 5;
        const int ca = a < 3 ? a : f + (a - 3), cb = b < 3 ? b : f + (b - 3);
        const int r = min(ca, cb), c = max(ca, cb);
        sPc[a][b] = pfix(praw(r, c), r, c);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double Pc[5][5];
        for (int a = 0; a < 5; a++)
            for (int b = 0; b < 5; b++) Pc[a][b] = sPc[a][b];
        gain_prologue(Xin, Pc, f, zr, zb, R, flags, g);
        if (!g.ok && count_status) atomicAdd(status, 1);
    }
    __syncthreads();
    for (int i = i0; i < n; i += stride) {
        double pc[5];  // P(i, c) for c in {0, 1, 2, f, f+1}
#pragma unroll
        for (int b = 0; b < 5; b++) {
            const int c = b < 3 ? b : f + (b - 3);
            double v;
            if (b < 3) {
                v = i <= c ? R3[(size_t)i * ld + c] : R3[(size_t)c * ld + i];
            } else if constexpr (SH) {
                v = colbuf[(size_t)(b - 3) * lda + i];
            } else {
                v = i <= c ? P[(size_t)i * ld + c] : P[(size_t)c * ld + i];
            }
            pc[b] = v;
        }
        for (int q = 0; q < kprev; q++) {
            const double* a0 = A + (size_t)2 * q * lda;
            const double* a1 = a0 + lda;
            const double a0i = a0[i], a1i = a1[i];
#pragma unroll
            for (int b = 0; b < 5; b++) {
                const int c = b < 3 ? b : f + (b - 3);
                pc[b] = pc[b] - rank2_term(a0i, a1i, a0[c], a1[c]);
            }
        }
        double pht[2];
        for (int k = 0; k < 2; k++)
            pht[k] = (((pc[0] * g.H[k][0] + pc[1] * g.H[k][1]) + pc[2] * g.H[k][2]) + pc[3] * g.H[k][3]) + pc[4] * g.H[k][4];
        const double w1_0 = pht[0] * g.G[0][0] + pht[1] * g.G[1][0];
        const double w1_1 = pht[0] * g.G[0][1] + pht[1] * g.G[1][1];
        const double w_0 = w1_0 * g.G[0][0] + w1_1 * g.G[0][1];
        const double w_1 = w1_0 * g.G[1][0] + w1_1 * g.G[1][1];
        Xout[i] = Xin[i] + (w_0 * g.V[0] + w_1 * g.V[1]);
        Aout[i] = w1_0;
        Aout[lda + i] = w1_1;
    }
}

template <bool SH>
__global__ void __launch_bounds__(256) k_gain_single(const double* Xin, double* Xout, const double* __restrict__ P,
                                                     const double* __restrict__ R3,
                                                     const double* __restrict__ colbuf, size_t ld, int n, double zr,
                                                     double zb, int idf, double r00, double r10, double r01,
                                                     double r11, unsigned flags, double* A, size_t lda,
                                                     int* __restrict__ status, const int* __restrict__ idf_dev,
                                                     int kprev) {
    __shared__ GainSmall g;
    __shared__ double sPc[5][5];
    if (idf_dev != nullptr) idf = *idf_dev;  // fused scan: the association index stays on the device
    const double R[4] = {r00, r10, r01, r11};
    gain_body<SH>(Xin, Xout, P, R3, colbuf, ld, n, zr, zb, idf, R, flags, A, lda, status, kprev,
                  blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, blockIdx.x == 0, g, sPc);
}

// Sharded column exchange: every rank contributes the entries of columns cols[k] (k < ncols) of
// the symmetric P that it stores — P[i][c] for owned rows i <= c, and P[c][i] (i > c) if it owns
// row c — zeros elsewhere; an all-reduce(sum) then yields the complete columns on every rank
// (x + 0 is exact, so the result is bit-identical to a single-GPU read).
struct ColList {
    int c[kMaxRank];
    int n;
};
// idf_dev (nullable): fused scan — column k belongs to observation k/2 whose 1-based landmark index is
// read from device memory (0 = no landmark passed the gate: the column is not needed, zeros are sent).
__global__ void __launch_bounds__(256) k_col_pack(const double* __restrict__ P, size_t ld, int n, ColList cl,
                                                  double* __restrict__ colbuf, size_t lda, Shard sh,
                                                  const int* __restrict__ idf_dev) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (i >= n) return;
    int c = cl.c[k];
    if (idf_dev != nullptr) {
        const int j = idf_dev[k >> 1];
        c = j > 0 ? 3 + 2 * (j - 1) + (k & 1) : -1;
    }
    double v = 0.0;
    if (c < 0) {
    } else if (i <= c) {
        if (shard_owns(sh, i)) v = P[shard_lrow(sh, i) * ld + c];
    } else {
        if (shard_owns(sh, c)) v = P[shard_lrow(sh, c) * ld + i];
    }
    colbuf[(size_t)k * lda + i] = v;
}

// Replicas of rows 0..2 (ranks != 0) apply the same rank-r update as k_cov_update applies to the
// authoritative rows on rank 0 — same operations, same order, bit-identical result.
__global__ void __launch_bounds__(256) k_rows012_update(double* __restrict__ R3, size_t ld, int n,
                                                        const double* __restrict__ A, size_t lda, int r,
                                                        double diag_eps) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= n || j < i) return;
    double s = A[i] * A[j];  // the operations of k_cov_update / k_cov_update_multi, in their order
    for (int k = 1; k < r; k++) s += A[(size_t)k * lda + i] * A[(size_t)k * lda + j];
    double o = R3[(size_t)i * ld + j] - s;
    if (j == i) o += diag_eps;
    R3[(size_t)i * ld + j] = o;
}

// Replicated cache of the landmarks' 2x2 diagonal blocks for the sharded gate: pack owned entries,
// zeros elsewhere, all-reduce.  D[0][j] = P_ff, D[1][j] = P_f,f+1, D[2][j] = P_f+1,f+1 (j 0-based).
__global__ void __launch_bounds__(256) k_diag_pack(const double* __restrict__ P, size_t ld, int nf,
                                                   double* __restrict__ D, int dcap, Shard sh) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nf) return;
    const int f = 3 + 2 * j;
    double a = 0.0, b = 0.0, c = 0.0;
    if (shard_owns(sh, f)) {
        a = P[shard_lrow(sh, f) * ld + f];
        b = P[shard_lrow(sh, f) * ld + f + 1];
    }
    if (shard_owns(sh, f + 1)) c = P[shard_lrow(sh, f + 1) * ld + f + 1];
    D[j] = a;
    D[(size_t)dcap + j] = b;
    D[2 * (size_t)dcap + j] = c;
}

// General-rank variant (r <= 64, any r) for the joint update on small maps; FP64 FMA, panels in
// shared memory.  The DMMA kernel in ekf_dmma.cu replaces it for large maps.
template <int T>
__global__ void __launch_bounds__(256) k_cov_update_rank(double* __restrict__ P, size_t ld, int n,
                                                         const double* __restrict__ A, size_t lda, int r, int nt,
                                                         Shard sh) {
    constexpr int CP = T / 2, RG = 256 / CP, RPT = T / RG;
    extern __shared__ double smem[];
    double* sAr = smem;          // [r][T]
    double* sAc = smem + r * T;  // [r][T]
    int tr, tc;
    shard_tile(blockIdx.x, nt, sh, tr, tc);
    const int i0 = tr * T, j0 = tc * T;
    for (int idx = threadIdx.x; idx < r * T; idx += 256) {
        const int k = idx / T, ii = idx % T;
        sAr[idx] = (i0 + ii < n) ? A[(size_t)k * lda + i0 + ii] : 0.0;
        sAc[idx] = (j0 + ii < n) ? A[(size_t)k * lda + j0 + ii] : 0.0;
    }
    __syncthreads();
    const int cp = threadIdx.x % CP, rg = threadIdx.x / CP;
    const int j = j0 + 2 * cp;
    if (j >= n) return;
    double s0[RPT], s1[RPT];
#pragma unroll
    for (int b = 0; b < RPT; b++) s0[b] = s1[b] = 0.0;
    for (int k = 0; k < r; k++) {
        const double a0 = sAc[k * T + 2 * cp], a1 = sAc[k * T + 2 * cp + 1];
#pragma unroll
        for (int b = 0; b < RPT; b++) {
            const double ai = sAr[k * T + rg + b * RG];
            s0[b] += ai * a0;
            s1[b] += ai * a1;
        }
    }
#pragma unroll
    for (int b = 0; b < RPT; b++) {
        const int i = i0 + rg + b * RG;
        if (i < n && j + 1 >= i) {
            double* p = P + shard_lrow(sh, i) * ld + j;
            double2 o = ld128(p);
            if (j >= i) o.x -= s0[b];
            if (j + 1 < n) o.y -= s1[b];
            st128(p, o);
        }
    }
}

// EKF.cpp:28-91 addOneNewFeature without the copy/resize: the two new columns (rows
// 0..len-1) and the new 2x2 diagonal block are written inside the pre-allocated P.
//   P[i][len+k] = (Gv * P[0:3, i])_k ;  P[len:len+2, len:len+2] = Gv Pvv Gv^T + Gz R Gz^T
// Sharded: every rank writes the rows it owns; rows 0..2 also go to the replicated panel R3.
// D (nullable): replicated diagonal-block cache — the new landmark's 2x2 block is entered directly.
__global__ void __launch_bounds__(256) k_augment(double* __restrict__ X, double* __restrict__ P,
                                                 double* __restrict__ R3, size_t ld, int len, double r, double b,
                                                 double r00, double r10, double r01, double r11, Shard sh,
                                                 double* __restrict__ D, int dcap) {
    const double phi = X[2];
    const double s = sin(phi + b), c = cos(phi + b);
    const double g02 = -r * s, g12 = r * c;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < len) {
        const double p0 = psym(R3, ld, 0, i), p1 = psym(R3, ld, 1, i), p2 = psym(R3, ld, 2, i);
        const double v0 = p0 + g02 * p2, v1 = p1 + g12 * p2;
        if (shard_owns(sh, i)) {
            P[shard_lrow(sh, i) * ld + len] = v0;
            P[shard_lrow(sh, i) * ld + len + 1] = v1;
        }
        if (i < 3 && R3 != P) {
            R3[(size_t)i * ld + len] = v0;
            R3[(size_t)i * ld + len + 1] = v1;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        X[len] = X[0] + r * c;  // only X[2], X[0], X[1] are read by other threads: no hazard
        X[len + 1] = X[1] + r * s;
        double Pv[3][3];
        for (int a = 0; a < 3; a++)
            for (int d = 0; d < 3; d++) Pv[a][d] = psym(R3, ld, a, d);
        const double Gv[2][3] = {{1, 0, g02}, {0, 1, g12}};
        const double Gz[2][2] = {{c, -r * s}, {s, r * c}};
        const double R[2][2] = {{r00, r01}, {r10, r11}};
        double GP[2][3], GR[2][2];
        for (int a = 0; a < 2; a++)
            for (int d = 0; d < 3; d++) {
                double acc = 0;
                for (int k = 0; k < 3; k++) acc += Gv[a][k] * Pv[k][d];
                GP[a][d] = acc;
            }
        for (int a = 0; a < 2; a++)
            for (int d = 0; d < 2; d++) {
                double acc = 0;
                for (int k = 0; k < 2; k++) acc += Gz[a][k] * R[k][d];
                GR[a][d] = acc;
            }
        for (int a = 0; a < 2; a++)
            for (int d = a; d < 2; d++) {
                double x = 0, y = 0;
                for (int k = 0; k < 3; k++) x += GP[a][k] * Gv[d][k];
                for (int k = 0; k < 2; k++) y += GR[a][k] * Gz[d][k];
                if (shard_owns(sh, len + a)) P[shard_lrow(sh, len + a) * ld + len + d] = x + y;
                if (D != nullptr) D[(size_t)(a + d) * dcap + (len - 3) / 2] = x + y;
            }
    }
}
// ---------------------------------------------------------------- joint (batch) update ----

struct ObsPack {  // kernel-parameter transport of one scan (no H2D copy)
    double z[2 * CSLAM_MAX_OBS];
    int idf[CSLAM_MAX_OBS];
    int m;
    double R[4];
};

// EKF.cpp:108-121: every observation linearised at the SAME prior X.
__global__ void k_batch_prep(const double* __restrict__ X, ObsPack ob, BatchSmall* __restrict__ sm) {
    const int k = threadIdx.x;
    if (k >= ob.m) return;
    const int f = 3 + 2 * (ob.idf[k] - 1);
    const ObsLin o = observe_lin(X[0], X[1], X[2], X[f], X[f + 1]);
    for (int a = 0; a < 2; a++) {
        for (int c = 0; c < 3; c++) sm->hu[k][a][c] = o.hu[a][c];
        sm->lu[k][a][0] = o.lu[a][0];
        sm->lu[k][a][1] = o.lu[a][1];
    }
    sm->f[k] = f;
    sm->V[2 * k] = ob.z[2 * k] - o.zr;
    sm->V[2 * k + 1] = pi2pi(ob.z[2 * k + 1] - o.zb);
}

// slam.h:243 PHT = P H^T with the stacked sparse H: 3 + 2m columns of P per state row.
// SH: the landmark columns come from the all-reduced exchange buffer (2k, 2k+1), rows 0..2 from R3.
template <bool SH>
__global__ void __launch_bounds__(128) k_batch_pht(const double* __restrict__ P, const double* __restrict__ R3,
                                                   const double* __restrict__ colbuf, size_t ld, int n, int m,
                                                   const BatchSmall* __restrict__ sm, double* __restrict__ PHT,
                                                   size_t lda) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (i >= n) return;
    const int f = sm->f[k];
    const double p0 = psym(R3, ld, i < 3 ? i : 0, i < 3 ? 0 : i);
    const double p1 = i < 3 ? psym(R3, ld, i, 1) : R3[ld + i];
    const double p2 = i < 3 ? psym(R3, ld, i, 2) : R3[2 * ld + i];
    double p3, p4;
    if constexpr (SH) {
        p3 = colbuf[(size_t)(2 * k) * lda + i];
        p4 = colbuf[(size_t)(2 * k + 1) * lda + i];
    } else {
        p3 = psym(P, ld, i, f);
        p4 = psym(P, ld, i, f + 1);
    }
    for (int a = 0; a < 2; a++)
        PHT[(size_t)(2 * k + a) * lda + i] =
            (((p0 * sm->hu[k][a][0] + p1 * sm->hu[k][a][1]) + p2 * sm->hu[k][a][2]) + p3 * sm->lu[k][a][0]) +
            p4 * sm->lu[k][a][1];
}

// slam.h:244-255: S = H PHT + RR, symmetrise, lower Cholesky, L^-1, finiteness check,
// G = L^-1 (literal) or L^-T (Q1), u = G G^T V.  One CTA, r <= 64, everything in shared memory.
// chol_smem: 2 * kMaxRank * (kMaxRank + 1) doubles of shared memory; tvec: kMaxRank doubles; okp: one int.
// Any block size (loops stride by blockDim.x); shared by k_batch_chol and the single-CTA observation step.
__device__ void batch_chol_body(const double* PHT, size_t lda, int m, double r00, double r10, double r01, double r11,
                                unsigned flags, BatchSmall* sm, int* status, double* chol_smem, double* tvec,
                                int* okp) {
    const int r = 2 * m;
    double(*S)[kMaxRank + 1] = reinterpret_cast<double(*)[kMaxRank + 1]>(chol_smem);
    double(*Li)[kMaxRank + 1] = reinterpret_cast<double(*)[kMaxRank + 1]>(chol_smem + kMaxRank * (kMaxRank + 1));
    int& ok = *okp;
    const double R[2][2] = {{r00, r01}, {r10, r11}};
    for (int idx = threadIdx.x; idx < r * r; idx += blockDim.x) {
        const int a = idx / r, b = idx % r;
        const int k = a / 2, al = a % 2;
        const int f = sm->f[k];
        double s = 0;
        for (int c = 0; c < 3; c++) s += sm->hu[k][al][c] * PHT[(size_t)b * lda + c];
        s += sm->lu[k][al][0] * PHT[(size_t)b * lda + f];
        s += sm->lu[k][al][1] * PHT[(size_t)b * lda + f + 1];
        if (a / 2 == b / 2) s += R[al][b % 2];
        S[a][b] = s;
    }
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    for (int idx = threadIdx.x; idx < r * r; idx += blockDim.x) {
        const int a = idx / r, b = idx % r;
        if (a > b) {
            const double v = (S[a][b] + S[b][a]) * 0.5;
            Li[a][b] = v;  // temp
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < r * r; idx += blockDim.x) {
        const int a = idx / r, b = idx % r;
        if (a > b) S[a][b] = S[b][a] = Li[a][b];
    }
    __syncthreads();
    // right-looking Cholesky on the lower triangle of S (in place)
    for (int j = 0; j < r; j++) {
        if (threadIdx.x == 0) {
            const double d = S[j][j];
            if (!(d > 0.0)) ok = 0;
            S[j][j] = sqrt(d);
        }
        __syncthreads();
        if (!ok) break;
        const double ljj = S[j][j];
        for (int i = j + 1 + threadIdx.x; i < r; i += blockDim.x) S[i][j] = S[i][j] / ljj;
        __syncthreads();
        for (int idx = threadIdx.x; idx < (r - j - 1) * (r - j - 1); idx += blockDim.x) {
            const int i = j + 1 + idx / (r - j - 1), c = j + 1 + idx % (r - j - 1);
            if (c <= i) S[i][c] -= S[i][j] * S[c][j];
        }
        __syncthreads();
    }
    // Li = L^-1 by forward substitution, one thread per column
    if (ok) {
        for (int c = threadIdx.x; c < r; c += blockDim.x) {
            for (int i = 0; i < r; i++) {
                if (i < c) { Li[i][c] = 0.0; continue; }
                double s = (i == c) ? 1.0 : 0.0;
                for (int k = c; k < i; k++) s -= S[i][k] * Li[k][c];
                Li[i][c] = s / S[i][i];
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && ok) {
        for (int i = 0; i < r && ok; i++)
            for (int c = 0; c <= i; c++)
                if (!isfinite(Li[i][c])) { ok = 0; break; }
    }
    __syncthreads();
    if (!ok && threadIdx.x == 0) atomicAdd(status, 1);
    // G and u
    for (int idx = threadIdx.x; idx < r * r; idx += blockDim.x) {
        const int a = idx / r, b = idx % r;
        double g = 0.0;
        if (ok) g = (flags & CSLAM_FLAG_Q1_METRIC_S) ? Li[b][a] : Li[a][b];
        sm->G[a * r + b] = g;
    }
    __syncthreads();
    // t = G^T V ; u = G t
    for (int l = threadIdx.x; l < r; l += blockDim.x) {
        double s = 0;
        for (int k = 0; k < r; k++) s += sm->G[k * r + l] * sm->V[k];
        tvec[l] = s;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < r; k += blockDim.x) {
        double s = 0;
        for (int l = 0; l < r; l++) s += sm->G[k * r + l] * tvec[l];
        sm->u[k] = s;
    }
}
__global__ void __launch_bounds__(256) k_batch_chol(const double* __restrict__ PHT, size_t lda, int m, double r00,
                                                    double r10, double r01, double r11, unsigned flags,
                                                    BatchSmall* __restrict__ sm, int* __restrict__ status) {
    extern __shared__ double chol_smem[];
    __shared__ double tvec[kMaxRank];
    __shared__ int ok;
    batch_chol_body(PHT, lda, m, r00, r10, r01, r11, flags, sm, status, chol_smem, tvec, &ok);
}

// slam.h:257-259: W1 = PHT G (panel A), Xout = Xin + PHT u  (= X + W1 G^T V).
// grid.y selects a chunk of 8 output ranks so accumulators stay in registers.
__global__ void __launch_bounds__(128) k_batch_w1(const double* __restrict__ PHT, size_t lda, int n, int r,
                                                  const BatchSmall* __restrict__ sm, const double* __restrict__ Xin,
                                                  double* __restrict__ Xout, double* __restrict__ A) {
    __shared__ double sG[kMaxRank][8];
    __shared__ double su[kMaxRank];
    const int l0 = blockIdx.y * 8;
    for (int idx = threadIdx.x; idx < r * 8; idx += blockDim.x) {
        const int k = idx / 8, l = idx % 8;
        sG[k][l] = (l0 + l < r) ? sm->G[k * r + l0 + l] : 0.0;
    }
    for (int k = threadIdx.x; k < r; k += blockDim.x) su[k] = sm->u[k];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double dx = 0;
    for (int k = 0; k < r; k++) {
        const double p = PHT[(size_t)k * lda + i];
#pragma unroll
        for (int l = 0; l < 8; l++) acc[l] += p * sG[k][l];
        dx += p * su[k];
    }
#pragma unroll
    for (int l = 0; l < 8; l++)
        if (l0 + l < r) A[(size_t)(l0 + l) * lda + i] = acc[l];
    if (blockIdx.y == 0) Xout[i] = Xin[i] + dx;
}

// Small maps (the reference's own 30-landmark world, n <= kSmallN): the whole observation step of
// test/main.cpp:186-189 — update(X, P, ZF, RE, IDF, batch = true) (EKF.cpp:93-129 -> slam.h:235-266) followed by
// augment(X, P, ZN, RE) (EKF.cpp:9-91) — in ONE launch of ONE CTA (the control steps are
// k_control_steps).  At n = 63 each of the 5 + mn kernels of the separate calls is pure launch latency; here the
// phases are separated by __syncthreads.  Every phase performs the operations of the kernel it replaces in the
// same order (k_batch_prep, k_batch_pht, k_batch_chol, k_batch_w1, k_cov_update_rank, k_augment), so the result
// is bit-identical to the separate calls.  X is NOT __restrict__ (threads communicate through it).
struct StepPack {
    ObsPack ob;                              // associated observations (may be empty)
    double zn[2 * CSLAM_MAX_BATCH_OBS];      // new landmarks (range, bearing)
    int mn;
};
__global__ void __launch_bounds__(1024) k_observe_step_small(const double* Xin, double* Xout, double* P, size_t ld, int n,
                                                             StepPack sp, unsigned flags, BatchSmall* sm, double* PHT,
                                                             double* A, size_t lda, int* status) {
    extern __shared__ double chol_smem[];
    __shared__ double tvec[kMaxRank];
    __shared__ int ok;
    const int tid = threadIdx.x, T = blockDim.x;
    const int m = sp.ob.m, r = 2 * m;
    double* X = const_cast<double*>(Xin);  // with no associated observation the state stays in its buffer
    if (m > 0) {
        X = Xout;
        // ---- k_batch_prep
        if (tid < m) {
            const int f = 3 + 2 * (sp.ob.idf[tid] - 1);
            const ObsLin o = observe_lin(Xin[0], Xin[1], Xin[2], Xin[f], Xin[f + 1]);
            for (int a = 0; a < 2; a++) {
                for (int c = 0; c < 3; c++) sm->hu[tid][a][c] = o.hu[a][c];
                sm->lu[tid][a][0] = o.lu[a][0];
                sm->lu[tid][a][1] = o.lu[a][1];
            }
            sm->f[tid] = f;
            sm->V[2 * tid] = sp.ob.z[2 * tid] - o.zr;
            sm->V[2 * tid + 1] = pi2pi(sp.ob.z[2 * tid + 1] - o.zb);
        }
        __syncthreads();
        // ---- k_batch_pht<false>
        for (int idx = tid; idx < m * n; idx += T) {
            const int k = idx / n, i = idx % n;
            const int f = sm->f[k];
            const double p0 = psym(P, ld, i < 3 ? i : 0, i < 3 ? 0 : i);
            const double p1 = i < 3 ? psym(P, ld, i, 1) : P[ld + i];
            const double p2 = i < 3 ? psym(P, ld, i, 2) : P[2 * ld + i];
            const double p3 = psym(P, ld, i, f), p4 = psym(P, ld, i, f + 1);
            for (int a = 0; a < 2; a++)
                PHT[(size_t)(2 * k + a) * lda + i] =
                    (((p0 * sm->hu[k][a][0] + p1 * sm->hu[k][a][1]) + p2 * sm->hu[k][a][2]) + p3 * sm->lu[k][a][0]) +
                    p4 * sm->lu[k][a][1];
        }
        __syncthreads();
        // ---- k_batch_chol
        batch_chol_body(PHT, lda, m, sp.ob.R[0], sp.ob.R[1], sp.ob.R[2], sp.ob.R[3], flags, sm, status, chol_smem, tvec, &ok);
        __syncthreads();
        // ---- k_batch_w1: W1 = PHT G (ascending k from 0.0), Xout = Xin + PHT u
        for (int i = tid; i < n; i += T) {
            double dx = 0;
            for (int k = 0; k < r; k++) dx += PHT[(size_t)k * lda + i] * sm->u[k];
            for (int l = 0; l < r; l++) {
                double acc = 0;
                for (int k = 0; k < r; k++) acc += PHT[(size_t)k * lda + i] * sm->G[k * r + l];
                A[(size_t)l * lda + i] = acc;
            }
            Xout[i] = Xin[i] + dx;
        }
        __syncthreads();
        // ---- k_cov_update_rank: P(i, j) -= sum_k A[k][i] A[k][j] over the upper triangle
        for (int idx = tid; idx < n * n; idx += T) {
            const int i = idx / n, j = idx % n;
            if (j < i) continue;
            double s0 = 0.0;
            for (int k = 0; k < r; k++) s0 += A[(size_t)k * lda + i] * A[(size_t)k * lda + j];
            P[(size_t)i * ld + j] -= s0;
        }
        __syncthreads();
    }
    // ---- k_augment, one new landmark after the other
    for (int q = 0; q < sp.mn; q++) {
        const int len = n + 2 * q;
        const double rr = sp.zn[2 * q], b = sp.zn[2 * q + 1];
        const double phi = X[2];
        const double sn = sin(phi + b), cs = cos(phi + b);
        const double g02 = -rr * sn, g12 = rr * cs;
        for (int i = tid; i < len; i += T) {
            const double p0 = psym(P, ld, 0, i), p1 = psym(P, ld, 1, i), p2 = psym(P, ld, 2, i);
            P[(size_t)i * ld + len] = p0 + g02 * p2;
            P[(size_t)i * ld + len + 1] = p1 + g12 * p2;
        }
        if (tid == 0) {
            X[len] = X[0] + rr * cs;
            X[len + 1] = X[1] + rr * sn;
            double Pv[3][3];
            for (int a = 0; a < 3; a++)
                for (int d = 0; d < 3; d++) Pv[a][d] = psym(P, ld, a, d);
            const double Gv[2][3] = {{1, 0, g02}, {0, 1, g12}};
            const double Gz[2][2] = {{cs, -rr * sn}, {sn, rr * cs}};
            const double R[2][2] = {{sp.ob.R[0], sp.ob.R[2]}, {sp.ob.R[1], sp.ob.R[3]}};
            double GP[2][3], GR[2][2];
            for (int a = 0; a < 2; a++)
                for (int d = 0; d < 3; d++) {
                    double acc = 0;
                    for (int k = 0; k < 3; k++) acc += Gv[a][k] * Pv[k][d];
                    GP[a][d] = acc;
                }
            for (int a = 0; a < 2; a++)
                for (int d = 0; d < 2; d++) {
                    double acc = 0;
                    for (int k = 0; k < 2; k++) acc += Gz[a][k] * R[k][d];
                    GR[a][d] = acc;
                }
            for (int a = 0; a < 2; a++)
                for (int d = a; d < 2; d++) {
                    double x = 0, y = 0;
                    for (int k = 0; k < 3; k++) x += GP[a][k] * Gv[d][k];
                    for (int k = 0; k < 2; k++) y += GR[a][k] * Gz[d][k];
                    P[(size_t)(len + a) * ld + len + d] = x + y;
                }
        }
        __syncthreads();
    }
}

// gating kernel lives in gate.cu (separate TU, compiled with -fmad=false)
int launch_gate(const double* X, const double* P, const double* R3, const double* D, int dcap, size_t ld, int nf,
                const double* Z, int m, const double R[4], double gate1, double gate2, double* part_nd,
                double* part_out, int* part_j, unsigned* ticket, int* jbest, double* nbest, double* outer,
                unsigned long long* assoc_count, cudaStream_t stream, int final_stage = 1, int* nblocks_out = nullptr);

// ------------------------------------------------------------------------ accessors ----
// sharded: every rank fills the entries it stores (zeros elsewhere) and the block is all-reduced.
// Rows 0..2 are read from R3 (rank 0's copy; on a lazy handle the big array's rows 0..2 are not maintained).
__global__ void k_gather_block(const double* __restrict__ P, const double* __restrict__ R3, size_t ld, int r0, int c0,
                               int nr, int nc, double* __restrict__ out, Shard sh) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)nr * nc) return;
    const int i = r0 + (int)(idx / nc), j = c0 + (int)(idx % nc);
    const int lo = i <= j ? i : j, hi = i <= j ? j : i;
    if (lo < 3)
        out[idx] = sh.rank == 0 ? R3[(size_t)lo * ld + hi] : 0.0;
    else
        out[idx] = shard_owns(sh, lo) ? P[shard_lrow(sh, lo) * ld + hi] : 0.0;
}
__global__ void k_scatter_upper(double* __restrict__ P, double* __restrict__ R3, size_t ld, int n,
                                const double* __restrict__ in, Shard sh) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * n) return;
    const int i = (int)(idx / n), j = (int)(idx % n);
    if (j < i) return;
    if (shard_owns(sh, i)) P[shard_lrow(sh, i) * ld + j] = in[idx];
    if (i < 3 && R3 != P) R3[(size_t)i * ld + j] = in[idx];
}

// defined in ekf_dmma.cu: tensor-core (FP64 DMMA) rank-r update for large maps
size_t dmma_panel_doubles(int n_cap);
int launch_cov_update_dmma(double* P, size_t ld, int n, const double* A, size_t lda, int r, Shard sh,
                           double* panels, int n_cap, int chunk, cudaStream_t stream, double* Pdst = nullptr,
                           double diag_eps = 0.0);

}  // namespace cslam
#include "ekf_lazy.cuh"
namespace cslam {

// ------------------------------------------------------------------------------------
// Host-side launch helpers
// ------------------------------------------------------------------------------------
static int allreduce_sum(cslam_ekf* h, double* buf, size_t count) {
    const NcclApi* api = nccl_api();
    if (!api) return CSLAM_ERR_NCCL;
    CSLAM_NCCL(api->AllReduce(buf, buf, count, ncclDouble, ncclSum, h->comm, h->stream));
    return CSLAM_OK;
}

// rows 0..2 on the replicas follow the authoritative update (same operations, same order)
static int replicas_follow(cslam_ekf* h, int r, double diag_eps) {
    if (h->sh.world == 1 || h->R3 == h->P) return CSLAM_OK;
    count_launch();
    k_rows012_update<<<dim3((h->n + 255) / 256, 3), 256, 0, h->stream>>>(h->R3, h->ld, h->n, h->A, h->lda, r,
                                                                        diag_eps);
    CSLAM_CUDA(cudaGetLastError());
    return CSLAM_OK;
}

// The replicated diagonal-block cache D of a sharded handle follows every FMA-path covariance pass the
// same way (same operations on its 3 entries per landmark), so the gate needs no pack + all-reduce
// per scan: terms = number of sequential updates in the pass, rows = panel rows per update (1: heading
// update slam.h:718 with its diagonal FLT_MIN, 2: landmark update).
__global__ void __launch_bounds__(256) k_diag_follow(double* __restrict__ D, int dcap, int nf,
                                                     const double* __restrict__ A, size_t lda, int terms, int rows,
                                                     double diag_eps, bool add_eps) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nf) return;
    const int f = 3 + 2 * j;
    double d00 = D[j], d01 = D[(size_t)dcap + j], d11 = D[2 * (size_t)dcap + j];
    for (int q = 0; q < terms; q++) {
        const double* a0 = A + (size_t)rows * q * lda;
        if (rows == 2) {
            const double* a1 = a0 + lda;
            d00 = d00 - rank2_term(a0[f], a1[f], a0[f], a1[f]);
            d01 = d01 - rank2_term(a0[f], a1[f], a0[f + 1], a1[f + 1]);
            d11 = d11 - rank2_term(a0[f + 1], a1[f + 1], a0[f + 1], a1[f + 1]);
        } else {
            d00 = d00 - a0[f] * a0[f];
            d01 = d01 - a0[f] * a0[f + 1];
            d11 = d11 - a0[f + 1] * a0[f + 1];
        }
        if (add_eps) {  // k_cov_update / k_cov_update_multi1 add diag_eps on the diagonal after every term
            d00 += diag_eps;  // (also when it is 0.0); k_cov_update_multi does not
            d11 += diag_eps;
        }
    }
    D[j] = d00;
    D[(size_t)dcap + j] = d01;
    D[2 * (size_t)dcap + j] = d11;
}
static int diag_follow(cslam_ekf* h, int terms, int rows, double diag_eps, bool add_eps) {
    if (h->sh.world == 1 || h->diag_dirty) return CSLAM_OK;  // no cache / cache stale anyway (next gate re-packs)
    const int nf = (h->n - 3) / 2;
    if (nf == 0) return CSLAM_OK;
    count_launch();
    k_diag_follow<<<(nf + 255) / 256, 256, 0, h->stream>>>(h->D, h->dcap, nf, h->A, h->lda, terms, rows, diag_eps,
                                                           add_eps);
    CSLAM_CUDA(cudaGetLastError());
    return CSLAM_OK;
}

// ... and the grouped form: g sequential rank-2 terms subtracted one after the other (k_cov_update_multi)
__global__ void __launch_bounds__(256) k_rows012_update_multi(double* __restrict__ R3, size_t ld, int n,
                                                              const double* __restrict__ A, size_t lda, int g) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= n || j < i) return;
    double o = R3[(size_t)i * ld + j];
    for (int q = 0; q < g; q++) {
        const double* a0 = A + (size_t)2 * q * lda;
        const double* a1 = a0 + lda;
        o = o - rank2_term(a0[i], a1[i], a0[j], a1[j]);
    }
    R3[(size_t)i * ld + j] = o;
}

template <int R>
static int launch_cov_update(cslam_ekf* h, double diag_eps, const int* live = nullptr) {
    const int n = h->n;
    {
        ProfScope prof(h);
        // tile / batch / occupancy / cache-policy chosen from tools/cov_variants.cu on a B200:
        // T=128, 4 x 16 B loads in flight per thread, 4 CTAs/SM, streaming (.cs) accesses -> 6.47 TB/s
        if (n >= 2048 || h->sh.world > 1) {
            const int nt = (n + 127) / 128;
            const long long tiles = shard_tile_count(nt, h->sh);
            if (tiles > 0) {
                count_launch();
                k_cov_update<R, 128, 4, 4, 1><<<(unsigned)tiles, 256, 0, h->stream>>>(h->P, h->ld, n, h->A, h->lda, nt,
                                                                                      diag_eps, h->sh, live);
            }
        } else {
            const int nt = (n + 63) / 64;
            count_launch();
            k_cov_update<R, 64, 8, 4, 0><<<(unsigned)shard_tile_count(nt, h->sh), 256, 0, h->stream>>>(
                h->P, h->ld, n, h->A, h->lda, nt, diag_eps, h->sh, live);
        }
        CSLAM_CUDA(cudaGetLastError());
    }
    if (int rc = diag_follow(h, 1, R, diag_eps, true)) return rc;
    return replicas_follow(h, R, diag_eps);
}

static int launch_cov_update_rank(cslam_ekf* h, int r) {
    const int n = h->n;
    if (h->lz.on) {
        // lazy handle (nothing pending, see cslam_ekf_update): tensor-core pass on the current array; rows 0..2
        // and the diagonal-block cache follow with the same FMA operations on every rank
        // Default: the round-1 kernel (ekf_dmma.cu: 128 x 32 tiles, 4-deep column-panel ring, covariance fragments by
        // per-thread cp.async) — 3.40 ms at N = 20k, 0.81 of the DMMA peak.  CSLAM_JOINT_TMA=1 selects the tensor-map
        // TMA variant (cov_tma.cu: 128 x 128 tiles, register-resident column fragments, direct stores), measured
        // SLOWER on a B200 (4.26 ms, DMMA pipe 65 %): with both 66 KB panels resident only a 2-deep ring fits and ncu
        // shows 13 % of the consumer time waiting for the single-buffered column panel (profiles/ncu_joint_tma_r02.csv).
        static const bool tma_kernel = getenv("CSLAM_JOINT_TMA") != nullptr;
        if (tma_kernel) {
            // tensor-map TMA loads, register-resident column fragments, direct stores (cov_tma.cu)
            ProfScope prof(h);
            if (int rc = launch_cov_update_tma_joint(h->lz.map[h->lz.stable], n, h->A, h->lda, r, h->sh, h->lz.num_sms,
                                                     lazy_P(h), h->ld, h->local_rows_cap, h->stream))
                return rc;
        } else {
            if (!h->dmma_panels) {
                const size_t bytes = dmma_panel_doubles(h->n_cap) * sizeof(double);
                CSLAM_CUDA(cudaMalloc(&h->dmma_panels, bytes));
                CSLAM_CUDA(cudaMemsetAsync(h->dmma_panels, 0, bytes, h->stream));
            }
            ProfScope prof(h);
            if (int rc = launch_cov_update_dmma(lazy_P(h), h->ld, n, h->A, h->lda, r, h->sh, h->dmma_panels, h->n_cap, 0,
                                                h->stream))
                return rc;
        }
        return lazy_follow(h, h->A, r, 0ULL);
    }
    h->diag_dirty = true;
    // large maps (and every sharded map): FP64 tensor-core kernel; small maps: plain FMA kernel
    if (n >= 1024 || h->sh.world > 1) {
        if (!h->dmma_panels) {
            const size_t bytes = dmma_panel_doubles(h->n_cap) * sizeof(double);
            CSLAM_CUDA(cudaMalloc(&h->dmma_panels, bytes));
            CSLAM_CUDA(cudaMemsetAsync(h->dmma_panels, 0, bytes, h->stream));
        }
        {
            ProfScope prof(h);
            if (int rc = launch_cov_update_dmma(h->P, h->ld, n, h->A, h->lda, r, h->sh, h->dmma_panels, h->n_cap, 0,
                                                h->stream))
                return rc;
        }
        if (h->sh.world > 1) {  // tensor-core rounding differs from the replicas' FMA order: re-broadcast rows 0..2
            const NcclApi* api = nccl_api();
            if (!api) return CSLAM_ERR_NCCL;
            CSLAM_NCCL(api->Broadcast(h->R3, h->R3, 3 * h->ld, ncclDouble, 0, h->comm, h->stream));
        }
        return CSLAM_OK;
    }
    ProfScope prof(h);
    const int nt = (n + 63) / 64;
    const size_t smem = (size_t)2 * r * 64 * sizeof(double);
    if (!h->attr_rank) {  // once per handle (device), not per call
        CSLAM_CUDA(cudaFuncSetAttribute(k_cov_update_rank<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        2 * kMaxRank * 64 * (int)sizeof(double)));
        h->attr_rank = true;
    }
    count_launch();
    k_cov_update_rank<64><<<(unsigned)shard_tile_count(nt, h->sh), 256, smem, h->stream>>>(h->P, h->ld, n, h->A,
                                                                                          h->lda, r, nt, h->sh);
    CSLAM_CUDA(cudaGetLastError());
    return CSLAM_OK;
}

// Covariance pass of a group of g sequential updates whose panels sit in rows 0..2g-1 of A.
static int launch_cov_update_multi(cslam_ekf* h, int g, const int* live) {
    if (g == 1) return launch_cov_update<2>(h, 0.0, live);
    const int n = h->n;
    {
    ProfScope prof(h);
    count_launch();
#define CSLAM_MULTI(M, MINB)                                                                                     \
    case M:                                                                                                      \
        if (n >= 2048 || h->sh.world > 1) {                                                                      \
            const int nt = (n + 127) / 128;                                                                      \
            if (shard_tile_count(nt, h->sh) == 0) break;                                                         \
            k_cov_update_multi<M, 128, 4, MINB, 1><<<(unsigned)shard_tile_count(nt, h->sh), 256, 0, h->stream>>>( \
                h->P, h->ld, n, h->A, h->lda, nt, h->sh, live);                                                  \
        } else {                                                                                                 \
            const int nt = (n + 63) / 64;                                                                        \
            k_cov_update_multi<M, 64, 4, 2, 0><<<(unsigned)shard_tile_count(nt, h->sh), 256, 0, h->stream>>>(  \
                h->P, h->ld, n, h->A, h->lda, nt, h->sh, live);                                                  \
        }                                                                                                        \
        break;
    switch (g) {
        // 4 resident CTAs / SM for every group size (64 registers: one update's column values at a
        // time); tools/cov_variants.cu on a B200, N=20k: M=2 2.27 ms, M=4 2.47 ms, M=8 3.14 ms
        CSLAM_MULTI(2, 4)
        CSLAM_MULTI(3, 4)
        CSLAM_MULTI(4, 4)
        CSLAM_MULTI(5, 4)
        CSLAM_MULTI(6, 4)
        CSLAM_MULTI(7, 4)
        CSLAM_MULTI(8, 4)
        default:
            set_last_error("launch_cov_update_multi: bad group size %d", g);
            return CSLAM_ERR_BAD_ARG;
    }
#undef CSLAM_MULTI
    }
    CSLAM_CUDA(cudaGetLastError());
    if (int rc = diag_follow(h, g, 2, 0.0, false)) return rc;
    if (h->sh.world > 1 && h->R3 != h->P) {  // replicas of rows 0..2 follow, same operations
        count_launch();
        k_rows012_update_multi<<<dim3((n + 255) / 256, 3), 256, 0, h->stream>>>(h->R3, h->ld, n, h->A, h->lda, g);
        CSLAM_CUDA(cudaGetLastError());
    }
    return CSLAM_OK;
}

// One pass for the k heading updates of k consecutive control steps (panel rows 0..k-1 of A); rows
// 0..2 were already brought up to date term by term (k_rows012_update), see k_cov_update_multi1.
static int launch_heading_multi(cslam_ekf* h, int k) {
    const int n = h->n;
    const int nt = (n + 127) / 128;
    const long long tiles = shard_tile_count(nt, h->sh);
    if (tiles == 0) return CSLAM_OK;
    ProfScope prof(h);
    count_launch();
#define CSLAM_H1(K)                                                                                          \
    case K:                                                                                                  \
        k_cov_update_multi1<K, 128, 4, 4, 1><<<(unsigned)tiles, 256, 0, h->stream>>>(h->P, h->ld, n, h->A, h->lda, nt, \
                                                                                     kFltMin, 3, h->sh);     \
        break;
    switch (k) {
        CSLAM_H1(1) CSLAM_H1(2) CSLAM_H1(3) CSLAM_H1(4) CSLAM_H1(5) CSLAM_H1(6) CSLAM_H1(7) CSLAM_H1(8)
        default:
            set_last_error("launch_heading_multi: bad group size %d", k);
            return CSLAM_ERR_BAD_ARG;
    }
#undef CSLAM_H1
    CSLAM_CUDA(cudaGetLastError());
    return CSLAM_OK;
}

constexpr int kSeqGroup = 8;  // sequential updates whose covariance passes are merged into one

// exchange the listed columns of P (sharded only)
static int exchange_columns(cslam_ekf* h, const ColList& cl, const int* idf_dev = nullptr) {
    count_launch();
    k_col_pack<<<dim3((h->n + 255) / 256, cl.n), 256, 0, h->stream>>>(h->P, h->ld, h->n, cl, h->colbuf, h->lda, h->sh,
                                                                      idf_dev);
    CSLAM_CUDA(cudaGetLastError());
    return allreduce_sum(h, h->colbuf, (size_t)cl.n * h->lda);
}

// singleUpdate (EKF.cpp:457-479): per observation a gain kernel (re-linearised at the X the previous
// observation produced, P seen through the pending rank-2 terms), per group of up to kSeqGroup
// observations ONE pass over the covariance.  idf_host / idf_dev: exactly one is non-null (idf_dev:
// fused scan, indices in device memory).  Sharded handles exchange the group's 2g columns of P first.
static int sequential_updates(cslam_ekf* h, const double* Z, const int32_t* idf_host, const int* idf_dev, int m,
                              const double R[4]) {
    const int n = h->n;
    const bool sharded = h->sh.world > 1;
    for (int base = 0; base < m; base += kSeqGroup) {
        const int g = std::min(kSeqGroup, m - base);
        if (sharded) {  // ONE exchange per group: the 2g observed columns of the not-yet-updated P
            ColList cl;
            cl.n = 2 * g;
            for (int k = 0; k < g; k++) {
                cl.c[2 * k] = idf_host ? 3 + 2 * (idf_host[base + k] - 1) : 0;
                cl.c[2 * k + 1] = cl.c[2 * k] + 1;
            }
            if (int rc = exchange_columns(h, cl, idf_dev ? idf_dev + base : nullptr)) return rc;
        }
        for (int k = 0; k < g; k++) {
            const int i = base + k;
            if (sharded) {
                count_launch();
                k_gain_single<true><<<(n + 255) / 256, 256, 0, h->stream>>>(
                    h->X[h->cur], h->X[h->cur ^ 1], h->P, h->R3, h->colbuf + (size_t)2 * k * h->lda, h->ld, n, Z[2 * i],
                    Z[2 * i + 1], idf_host ? idf_host[i] : 0, R[0], R[1], R[2], R[3], h->flags, h->A, h->lda, h->status,
                    idf_dev ? idf_dev + i : nullptr, k);
            } else {
                count_launch();
                k_gain_single<false><<<(n + 255) / 256, 256, 0, h->stream>>>(
                    h->X[h->cur], h->X[h->cur ^ 1], h->P, h->R3, nullptr, h->ld, n, Z[2 * i], Z[2 * i + 1],
                    idf_host ? idf_host[i] : 0, R[0], R[1], R[2], R[3], h->flags, h->A, h->lda, h->status,
                    idf_dev ? idf_dev + i : nullptr, k);
            }
            CSLAM_CUDA(cudaGetLastError());
            h->cur ^= 1;
        }
        if (int rc = launch_cov_update_multi(h, g, idf_dev ? idf_dev + base : nullptr)) return rc;
    }
    return CSLAM_OK;
}

static int check_handle(const cslam_ekf* h) {
    CSLAM_REQUIRE(h != nullptr, CSLAM_ERR_BAD_ARG, "null handle");
    CSLAM_CUDA(cudaSetDevice(h->device));
    return CSLAM_OK;
}

static int create_common(cslam_ekf_t** out, int capacity_landmarks, int device, unsigned flags, int rank, int world,
                         const void* nccl_id) {
    CSLAM_REQUIRE(out != nullptr, CSLAM_ERR_BAD_ARG, "out is null");
    CSLAM_REQUIRE(capacity_landmarks >= 0 && capacity_landmarks <= 500000, CSLAM_ERR_BAD_ARG,
                  "capacity_landmarks out of range");
    CSLAM_REQUIRE(world >= 1 && rank >= 0 && rank < world, CSLAM_ERR_BAD_ARG, "bad rank/world");
    CSLAM_REQUIRE(world == 1 || nccl_id != nullptr, CSLAM_ERR_BAD_ARG, "sharded handle needs an NCCL unique id");
    *out = nullptr;
    int count = 0;
    CSLAM_CUDA(cudaGetDeviceCount(&count));
    CSLAM_REQUIRE(device >= 0 && device < count, CSLAM_ERR_CUDA, "no such CUDA device (no CPU fallback exists)");
    CSLAM_CUDA(cudaSetDevice(device));
    cslam_ekf* h = new (std::nothrow) cslam_ekf();
    CSLAM_REQUIRE(h != nullptr, CSLAM_ERR_BAD_ARG, "out of host memory");
    h->device = device;
    h->flags = flags;
    h->cap_landmarks = capacity_landmarks;
    h->n_cap = 3 + 2 * capacity_landmarks;
    h->ld = ((size_t)h->n_cap + 1 + 15) / 16 * 16;  // >= n_cap + 1 so the last column pair stays in-row
    h->lda = h->ld;
    h->n = 3;
    h->sh = Shard{rank, world};
    // rows stored here: every owned 128-row tile of the capacity (all rows when world == 1)
    {
        const int ntr = (h->n_cap + kShardRows - 1) / kShardRows;
        int owned = 0;
        for (int tr = rank; tr < ntr; tr += world) owned++;
        h->local_rows_cap = world == 1 ? h->n_cap : std::max(owned, 1) * kShardRows;
    }
    auto fail = [&](int code) {
        cslam_ekf_destroy(h);
        return code;
    };
    const size_t pbytes = (size_t)h->local_rows_cap * h->ld * sizeof(double);
#define TRY(call)                                                                            \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            set_last_error("cslam_ekf_create: %s -> %s", #call, cudaGetErrorString(e__));    \
            return fail(CSLAM_ERR_CUDA);                                                     \
        }                                                                                    \
    } while (0)
    {
        // The handle's own stream gets the HIGHEST priority: on a deferred-pass handle the small kernels of the
        // per-scan chain must be dispatched into the resources a running covariance pass leaves free, ahead of that
        // pass's own pending CTAs (equal priorities are served in launch order: the chain would wait for the whole
        // grid of the tensor-core pass to be dispatched).  A caller stream (cslam_ekf_set_stream) should be created
        // with a higher priority than the default for the same reason; the pass stream has the lowest.
        int least = 0, greatest = 0;
        TRY(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        TRY(cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, greatest));
    }
    TRY(cudaEventCreateWithFlags(&h->scan_ev, cudaEventDisableTiming));
    h->own_stream = true;
    TRY(cudaMalloc(&h->X[0], h->ld * sizeof(double)));
    TRY(cudaMalloc(&h->X[1], h->ld * sizeof(double)));
    TRY(cudaMalloc(&h->P, pbytes));
    {  // deferred covariance passes (ekf_lazy.cuh): large and sharded maps; CSLAM_LAZY=0/1 overrides
        const char* e = getenv("CSLAM_LAZY");
        h->lz.on = e ? atoi(e) != 0 : (h->n_cap >= 2048 || world > 1);
    }
    const size_t a_rows = h->lz.on ? (size_t)std::max(kMaxRank, 2 * kLazyBankMax) : (size_t)kMaxRank;  // two banks when lazy
    TRY(cudaMalloc(&h->A, a_rows * h->lda * sizeof(double)));
    TRY(cudaMalloc(&h->PHT, (size_t)kMaxRank * h->lda * sizeof(double)));
    TRY(cudaMalloc(&h->small, sizeof(BatchSmall)));
    TRY(cudaMalloc(&h->status, sizeof(int)));
    TRY(cudaMalloc(&h->ticket, 4 * sizeof(unsigned)));
    h->gate.max_blocks = (capacity_landmarks + 127) / 128 + 1;  // k_gate runs 128-thread blocks
    TRY(cudaMalloc(&h->gate.part_nd, (size_t)h->gate.max_blocks * CSLAM_MAX_OBS * sizeof(double)));
    TRY(cudaMalloc(&h->gate.part_out, (size_t)h->gate.max_blocks * CSLAM_MAX_OBS * sizeof(double)));
    TRY(cudaMalloc(&h->gate.part_j, (size_t)h->gate.max_blocks * CSLAM_MAX_OBS * sizeof(int)));
    TRY(cudaMalloc(&h->gate.d_jbest, CSLAM_MAX_OBS * sizeof(int)));
    TRY(cudaMalloc(&h->gate.d_nbest, CSLAM_MAX_OBS * sizeof(double)));
    TRY(cudaMalloc(&h->gate.d_outer, CSLAM_MAX_OBS * sizeof(double)));
    TRY(cudaMalloc(&h->assoc_count, sizeof(unsigned long long)));
    TRY(cudaMemsetAsync(h->assoc_count, 0, sizeof(unsigned long long), h->stream));
    h->pinned_bytes = std::max<size_t>(h->ld * sizeof(double), 1 << 16);
    TRY(cudaMallocHost(&h->pinned, h->pinned_bytes));
    TRY(cudaMemsetAsync(h->X[0], 0, h->ld * sizeof(double), h->stream));
    TRY(cudaMemsetAsync(h->X[1], 0, h->ld * sizeof(double), h->stream));
    TRY(cudaMemsetAsync(h->P, 0, pbytes, h->stream));
    TRY(cudaMemsetAsync(h->A, 0, a_rows * h->lda * sizeof(double), h->stream));
    TRY(cudaMemsetAsync(h->PHT, 0, (size_t)kMaxRank * h->lda * sizeof(double), h->stream));
    TRY(cudaMemsetAsync(h->status, 0, sizeof(int), h->stream));
    TRY(cudaMemsetAsync(h->ticket, 0, 4 * sizeof(unsigned), h->stream));
    h->R3 = h->P;  // one GPU, or rank 0 of a sharded handle: rows 0..2 are the first rows of P
    if (world > 1 || h->lz.on) {
        if (rank != 0 || h->lz.on) {  // lazy handles keep rows 0..2 in their own panel on EVERY rank
            TRY(cudaMalloc(&h->R3, 3 * h->ld * sizeof(double)));
            TRY(cudaMemsetAsync(h->R3, 0, 3 * h->ld * sizeof(double), h->stream));
        }
        TRY(cudaMalloc(&h->colbuf, (size_t)kMaxRank * h->lda * sizeof(double)));
        h->dcap = std::max(capacity_landmarks, 1);
        TRY(cudaMalloc(&h->D, 3 * (size_t)h->dcap * sizeof(double)));
        TRY(cudaMemsetAsync(h->D, 0, 3 * (size_t)h->dcap * sizeof(double), h->stream));
    }
    if (h->lz.on) {
        LazyState& L = h->lz;
        L.Pbuf[0] = h->P;
        TRY(cudaDeviceGetAttribute(&L.num_sms, cudaDevAttrMultiProcessorCount, device));
        {
            int least = 0, greatest = 0;
            TRY(cudaDeviceGetStreamPriorityRange(&least, &greatest));
            TRY(cudaStreamCreateWithPriority(&L.pass_stream, cudaStreamNonBlocking, least));
        }
        TRY(cudaEventCreateWithFlags(&L.ev_chain, cudaEventDisableTiming));
        TRY(cudaEventCreateWithFlags(&L.ev_pass, cudaEventDisableTiming));
        if (const char* e = getenv("CSLAM_TMA_STAGES")) L.stages = atoi(e);
        if (const char* e = getenv("CSLAM_GAIN_FUSED")) L.fused_gains = atoi(e) != 0;
        if (const char* e = getenv("CSLAM_GATE_DEFER")) L.defer_gate_merge = atoi(e) != 0;
        if (getenv("CSLAM_KTRACE")) {
            TRY(cudaMalloc(&L.ktrace, (size_t)8 * LazyState::kTraceCap * sizeof(unsigned long long)));
            TRY(cudaMemsetAsync(L.ktrace, 0, (size_t)8 * LazyState::kTraceCap * sizeof(unsigned long long), h->stream));
        }
        TRY(cudaMalloc(&L.R3alt, 3 * h->ld * sizeof(double)));
        TRY(cudaMemsetAsync(L.R3alt, 0, 3 * h->ld * sizeof(double), h->stream));
        TRY(cudaMalloc(&L.hdr, sizeof(GroupHeader)));
        TRY(cudaMemsetAsync(L.hdr, 0, sizeof(GroupHeader), h->stream));
        TRY(cudaMalloc(&L.Dalt, 3 * (size_t)h->dcap * sizeof(double)));
        TRY(cudaMemsetAsync(L.Dalt, 0, 3 * (size_t)h->dcap * sizeof(double), h->stream));
        // ping-pong pair: the pass of bank k streams array a -> array b while the gains of the following scans
        // read array a, so the chain never waits for a running pass.  Costs a second covariance array; taken
        // when it fits comfortably (CSLAM_PINGPONG=0/1 overrides).
        size_t free_b = 0, total_b = 0;
        TRY(cudaMemGetInfo(&free_b, &total_b));
        const char* e = getenv("CSLAM_PINGPONG");
        L.pingpong = e ? atoi(e) != 0 : (free_b > pbytes + pbytes / 8 + ((size_t)2 << 30));
        if (L.pingpong) {
            TRY(cudaMalloc(&L.Pbuf[1], pbytes));
            TRY(cudaMemsetAsync(L.Pbuf[1], 0, pbytes, h->stream));
        }
        for (int b = 0; b < (L.pingpong ? 2 : 1); b++)
            if (make_cov_tensor_map(L.map[b], L.Pbuf[b], h->ld, (size_t)h->local_rows_cap) != CSLAM_OK)
                return fail(CSLAM_ERR_CUDA);
        if (!L.pingpong) memcpy(L.map[1], L.map[0], sizeof(L.map[0]));
        // Bank size: 64 rows (32 sequential updates per read + write of the covariance, tensor-core pass) where the
        // pass still is the longer side of a scan — one or two GPUs; 32 rows from four GPUs on, where the per-scan
        // chain (whose cost grows with the number of pending terms) would otherwise become the limit.  In place
        // (no ping-pong twin) the chain waits for every pass, so the bank that keeps the TMA pass is used.
        L.bank_rows = !L.pingpong ? kLazyBankTma : (world <= 2 ? 64 : 32);
        if (const char* eb = getenv("CSLAM_LAZY_BANK")) {
            const int v = atoi(eb);
            if (v >= 2 && v <= kLazyBankMax && (v % 2) == 0) L.bank_rows = v;
        }
        if (L.bank_rows > kLazyBankTma) {
            const size_t bytes = dmma_panel_doubles(h->n_cap) * sizeof(double);
            TRY(cudaMalloc(&L.pass_panels, bytes));
            TRY(cudaMemsetAsync(L.pass_panels, 0, bytes, h->stream));
        }
        if (world > 1) {  // snapshot buffers + flags of the peer-memory column exchange (mapped by cslam_ekf_ipc_*)
            const size_t xb = (size_t)2 * 2 * kSeqGroupLazyMax * h->lda * sizeof(double);
            // ... followed by the flagged-cell copy of the same buffers (16-byte cells, k_col_push_ll): one allocation,
            // one IPC handle
            L.xll_off = xb;
            TRY(cudaMalloc(&L.xbuf, 3 * xb));
            TRY(cudaMemsetAsync(L.xbuf, 0, 3 * xb, h->stream));
            TRY(cudaMalloc(&L.sig, 8 * sizeof(unsigned long long)));
            TRY(cudaMemsetAsync(L.sig, 0, 8 * sizeof(unsigned long long), h->stream));
            TRY(cudaMalloc(&L.push_ticket, sizeof(unsigned)));
            TRY(cudaMemsetAsync(L.push_ticket, 0, sizeof(unsigned), h->stream));
            L.peer_xbuf[rank] = L.xbuf;
            L.peer_sig[rank] = L.sig;
        }
    }
    if (world > 1) {
        const NcclApi* api = nccl_api();
        if (!api) return fail(CSLAM_ERR_NCCL);
        ncclUniqueId id;
        memcpy(&id, nccl_id, sizeof(id));
        ncclResult_t r = api->CommInitRank(&h->comm, world, id, rank);
        if (r != ncclSuccess) {
            set_last_error("ncclCommInitRank -> %s", api->GetErrorString(r));
            return fail(CSLAM_ERR_NCCL);
        }
    }
    TRY(cudaStreamSynchronize(h->stream));
#undef TRY
    *out = h;
    return CSLAM_OK;
}

}  // namespace cslam

using namespace cslam;

// ------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------
extern "C" {

int cslam_ekf_create(cslam_ekf_t** out, int capacity_landmarks, int device, unsigned flags) {
    CSLAM_NVTX_RANGE();
    return create_common(out, capacity_landmarks, device, flags, 0, 1, nullptr);
}

int cslam_ekf_create_sharded(cslam_ekf_t** out, int capacity_landmarks, int device, unsigned flags, int rank,
                             int world, const void* nccl_unique_id) {
    CSLAM_NVTX_RANGE();
    return create_common(out, capacity_landmarks, device, flags, rank, world, nccl_unique_id);
}

int cslam_ekf_destroy(cslam_ekf_t* h) {
    CSLAM_NVTX_RANGE();
    if (!h) return CSLAM_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->lz.pass_stream) {
        cudaStreamSynchronize(h->lz.pass_stream);
        cudaStreamDestroy(h->lz.pass_stream);
    }
    if (h->lz.ev_chain) cudaEventDestroy(h->lz.ev_chain);
    if (h->lz.ev_pass) cudaEventDestroy(h->lz.ev_pass);
    if (h->lz.peers_ready)
        for (int q = 0; q < h->sh.world; q++)
            if (q != h->sh.rank) {
                cudaIpcCloseMemHandle(h->lz.peer_xbuf[q]);
                cudaIpcCloseMemHandle(h->lz.peer_sig[q]);
            }
    cudaFree(h->lz.xbuf);
    cudaFree(h->lz.sig);
    cudaFree(h->lz.push_ticket);
    cudaFree(h->lz.Pbuf[1]);
    cudaFree(h->lz.R3alt);
    cudaFree(h->lz.Dalt);
    cudaFree(h->lz.pass_panels);
    cudaFree(h->lz.hdr);
    if (h->lz.ktrace) {  // diagnostics: dump the kernel timestamps
        const char* path = getenv("CSLAM_KTRACE");
        const int cnt = std::min(h->lz.kslot, LazyState::kTraceCap);
        std::vector<unsigned long long> t((size_t)8 * cnt);
        if (path && cnt > 0 &&
            cudaMemcpy(t.data(), h->lz.ktrace, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess) {
            const std::string file = std::string(path) + ".r" + std::to_string(h->sh.rank);
            if (FILE* f = fopen(file.c_str(), "ab")) {
                fwrite(t.data(), sizeof(unsigned long long), t.size(), f);
                fclose(f);
            }
        }
        cudaFree(h->lz.ktrace);
    }
    cudaFree(h->trace_dev);
    cudaFree(h->acc_dev);
    if (h->comm) {
        const NcclApi* api = nccl_api();
        if (api) api->CommDestroy(h->comm);
    }
    if (h->R3 && h->R3 != h->P) cudaFree(h->R3);
    cudaFree(h->colbuf); cudaFree(h->D);
    cudaFree(h->X[0]); cudaFree(h->X[1]); cudaFree(h->P); cudaFree(h->A); cudaFree(h->PHT);
    cudaFree(h->dmma_panels);
    cudaFree(h->small); cudaFree(h->status); cudaFree(h->ticket);
    cudaFree(h->gate.part_nd); cudaFree(h->gate.part_out); cudaFree(h->gate.part_j);
    cudaFree(h->gate.d_jbest); cudaFree(h->gate.d_nbest); cudaFree(h->gate.d_outer);
    cudaFree(h->assoc_count);
    if (h->pinned) cudaFreeHost(h->pinned);
    for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
    if (h->scan_ev) cudaEventDestroy(h->scan_ev);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return CSLAM_OK;
}

int cslam_ekf_set_stream(cslam_ekf_t* h, void* cuda_stream) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_handle(h)) return rc;
    CSLAM_CUDA(cudaStreamSynchronize(h->stream));
    if (h->own_stream) cudaStreamDestroy(h->stream);
    h->stream = static_cast<cudaStream_t>(cuda_stream);
    h->own_stream = false;
    return CSLAM_OK;
}

// Peer-memory column exchange of sharded handles: 2 cudaIpcMemHandle_t (snapshot buffers, flags) per rank.
int cslam_ekf_ipc_export(cslam_ekf_t* h, void* out128) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_handle(h)) return rc;
    CSLAM_REQUIRE(out128 != nullptr, CSLAM_ERR_BAD_ARG, "out is null");
    CSLAM_REQUIRE(h->sh.world > 1 && h->lz.on && h->lz.xbuf, CSLAM_ERR_UNSUPPORTED, "not a sharded lazy handle");
    cudaIpcMemHandle_t* hs = static_cast<cudaIpcMemHandle_t*>(out128);
    CSLAM_CUDA(cudaIpcGetMemHandle(&hs[0], h->lz.xbuf));
    CSLAM_CUDA(cudaIpcGetMemHandle(&hs[1], h->lz.sig));
    return CSLAM_OK;
}
// all: world x 128 bytes in rank order.  Without this call a sharded handle exchanges its columns through NCCL.
int cslam_ekf_ipc_import(cslam_ekf_t* h, const void* all, int world) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_handle(h)) return rc;
    CSLAM_REQUIRE(all != nullptr && world == h->sh.world && world <= 8, CSLAM_ERR_BAD_ARG, "bad argument");
    CSLAM_REQUIRE(h->lz.on && h->lz.xbuf, CSLAM_ERR_UNSUPPORTED, "not a sharded lazy handle");
    const cudaIpcMemHandle_t* hs = static_cast<const cudaIpcMemHandle_t*>(all);
    for (int q = 0; q < world; q++) {
        if (q == h->sh.rank) continue;
        void *px = nullptr, *ps = nullptr;
        CSLAM_CUDA(cudaIpcOpenMemHandle(&px, hs[2 * q], cudaIpcMemLazyEnablePeerAccess));
        CSLAM_CUDA(cudaIpcOpenMemHandle(&ps, hs[2 * q + 1], cudaIpcMemLazyEnablePeerAccess));
        h->lz.peer_xbuf[q] = static_cast<double*>(px);
        h->lz.peer_sig[q] = static_cast<unsigned long long*>(ps);
    }
    h->lz.peers_ready = true;
    return CSLAM_OK;
}

int cslam_ekf_flush(cslam_ekf_t* h) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_handle(h)) return rc;
    return lazy_flush_all(h);
}

int cslam_ekf_pass_count(cslam_ekf_t* h, unsigned long long* passes, int* pending_rows) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_handle(h)) return rc;
    if (passes) *passes = h->lz.passes;
    if (pending_rows) *pending_rows = h->lz.on ? h->lz.np : 0;
    return CSLAM_OK;
}

int cslam_ekf_sync(cslam_ekf_t* h, int* skipped_updates) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_handle(h)) return rc;
    if (int rc = lazy_flush_all(h)) return rc;
    if (skipped_updates) {
        CSLAM_CUDA(cudaMemcpyAsync(h->pinned, h->status, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CSLAM_CUDA(cudaStreamSynchronize(h->stream));
        *skipped_updates = *static_cast<int*>(h->pinned);
    } else {
        CSLAM_CUDA(cudaStreamSynchronize(h->stream));
    }
    return CSLAM_OK;
}

int cslam_ekf_n(const cslam_ekf_t* h) { return h ? h->n : -1; }
int cslam_ekf_num_landmarks(const cslam_ekf_t* h) { return h ? (h->n - 3) / 2 : -1; }
int cslam_ekf_capacity(const cslam_ekf_t* h) { return h ? h->cap_landmarks : -1; }

int cslam_ekf_predict(cslam_ekf_t* h, double v, double swa, const double Q[4], double wb, double dt) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_handle(h)) return rc;
    CSLAM_REQUIRE(Q != nullptr, CSLAM_ERR_BAD_ARG, "Q is null");
    const int n = h->n;
    int width = 0;
    if (n > 3) width = (h->flags & CSLAM_FLAG_Q2_FULL_WIDTH) ? n - 3 : n - 4;  // Q2, EKF.cpp:442
    const int blocks = std::max(1, (width + 255) / 256);
    // rows 0..2 only: on a sharded handle every rank predicts its replica (rank 0: the rows themselves)
    count_launch();
    k_predict<<<blocks, 256, 0, h->stream>>>(h->X[h->cur], h->R3, h->ld, n, v, swa, Q[0], Q[2], Q[1], Q[3], wb, dt,
                                             width, h->ticket);
    CSLAM_CUDA(cudaGetLastError());
    return CSLAM_OK;
}

int cslam_ekf_observe_heading(cslam_ekf_t* h, double phi, int use_heading) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_handle(h)) return rc;
    if (!use_heading) return CSLAM_OK;  // EKF.cpp:332-335
    if (h->lz.on) return lazy_heading(h, phi);
    const int n = h->n;
    const double sigma = 0.01F * kPi / 180.0F;  // EKF.cpp:337
    count_launch();
    k_heading_gain<<<(n + 255) / 256, 256, 0, h->stream>>>(h->X[h->cur], h->X[h->cur ^ 1], h->R3, h->ld, n, phi,
                                                           sigma * sigma, h->A);
    CSLAM_CUDA(cudaGetLastError());
    h->cur ^= 1;
    return launch_cov_update<1>(h, kFltMin);
}

// Sharded handles: (re)build the replicated diagonal-block cache the gate reads — pack owned entries +
// all-reduce.  Needed only after reset / augment / a tensor-core joint update; the FMA-path updates
// keep it current themselves (k_diag_follow).
static int refresh_diag_cache(cslam_ekf* h) {
    if ((h->sh.world == 1 && !h->lz.on) || !h->diag_dirty) return CSLAM_OK;
    if (int rc = lazy_flush_all(h)) return rc;
    const int nf = (h->n - 3) / 2;
    CSLAM_CUDA(cudaMemsetAsync(h->D, 0, 3 * (size_t)h->dcap * sizeof(double), h->stream));
    if (nf > 0) {
        count_launch();
        k_diag_pack<<<(nf + 255) / 256, 256, 0, h->stream>>>(lazy_P(h), h->ld, nf, h->D, h->dcap, h->sh);
        CSLAM_CUDA(cudaGetLastError());
    }
    if (h->sh.world > 1)
        if (int rc = allreduce_sum(h, h->D, 3 * (size_t)h->dcap)) return rc;
    h->diag_dirty = false;
    return CSLAM_OK;
}

// device scratch owned by the handle: grown when a call needs more, never allocated per call
static int ensure_scratch(cslam_ekf* h, double** buf, size_t* cap, size_t doubles) {
    if (*cap >= doubles) return CSLAM_OK;
    CSLAM_CUDA(cudaStreamSynchronize(h->stream));
    cudaFree(*buf);
    *buf = nullptr;
    *cap = 0;
    const size_t want = std::max<size_t>(doubles, 4096);
    CSLAM_CUDA(cudaMalloc(buf, want * sizeof(double)));
    *cap = want;
    return CSLAM_OK;
}

int cslam_ekf_control_steps(cslam_ekf_t* h, int k, const double* v, const double* swa, const double* phi,
                            int use_heading, const double Q[4], double wb, double dt, double* pose_trace) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_handle(h)) return rc;
    CSLAM_REQUIRE(k >= 0, CSLAM_ERR_BAD_ARG, "k < 0");
    if (k == 0) return CSLAM_OK;
    CSLAM_REQUIRE(v && swa && Q && (phi || !use_heading), CSLAM_ERR_BAD_ARG, "null argument");
    const double sigma = 0.01F * kPi / 180.0F;  // EKF.cpp:337
    double* trace_dev = nullptr;
    if (pose_trace) {
        if (int rc = ensure_scratch(h, &h->trace_dev, &h->trace_cap, (size_t)3 * k)) return rc;
        trace_dev = h->trace_dev;
    }
    int rc = CSLAM_OK;
    if (h->lz.on) {
        // lazy handle: per step the O(n) kernels only (predict on rows 0..2, heading gain, R3 / D follow); the
        // rank-1 terms join the pending bank and are applied by the next pass together with whatever else
        // is pending — a whole drive cycle of test/main.cpp streams the covariance once
        for (int st = 0; st < k && rc == CSLAM_OK; st++) {
            rc = cslam_ekf_predict(h, v[st], swa[st], Q, wb, dt);
            if (rc == CSLAM_OK && use_heading) rc = lazy_heading(h, phi[st]);
            if (rc == CSLAM_OK && trace_dev) {
                cudaError_t e = cudaMemcpyAsync(trace_dev + 3 * st, h->X[h->cur], 3 * sizeof(double),
                                                cudaMemcpyDeviceToDevice, h->stream);
                if (e != cudaSuccess) rc = CSLAM_ERR_CUDA;
            }
        }
        k = rc == CSLAM_OK ? k : 0;
    }
    for (int base = 0; base < k && rc == CSLAM_OK && !h->lz.on; base += kMaxControlSteps) {
        const int kc = std::min(kMaxControlSteps, k - base);
        const int n = h->n;
        int width = 0;
        if (n > 3) width = (h->flags & CSLAM_FLAG_Q2_FULL_WIDTH) ? n - 3 : n - 4;  // Q2, EKF.cpp:442
        if (n <= kSmallN && h->sh.world == 1) {
            ControlPack cp;
            memset(&cp, 0, sizeof(cp));
            for (int i = 0; i < kc; i++) {
                cp.v[i] = v[base + i];
                cp.swa[i] = swa[base + i];
                cp.phi[i] = phi ? phi[base + i] : 0.0;
            }
            cp.k = kc; cp.use_heading = use_heading ? 1 : 0; cp.width = width;
            cp.q00 = Q[0]; cp.q01 = Q[2]; cp.q10 = Q[1]; cp.q11 = Q[3];
            cp.wb = wb; cp.dt = dt; cp.r_heading = sigma * sigma;
            count_launch();
            k_control_steps<<<1, 1024, 0, h->stream>>>(h->X[h->cur], h->P, h->ld, n, cp,
                                                       trace_dev ? trace_dev + 3 * base : nullptr);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) {
                set_last_error("k_control_steps -> %s", cudaGetErrorString(e));
                rc = CSLAM_ERR_CUDA;
            }
        } else {
            // big or sharded maps: per step the O(n) kernels (predict, heading gain, rank-1 update of rows
            // 0..2 and of the replicated caches), then ONE pass over the rest of the covariance for up to
            // 8 heading updates (k_cov_update_multi1) instead of one pass per control step
            for (int g0 = 0; g0 < kc && rc == CSLAM_OK; g0 += kSeqGroup) {
                const int gk = std::min(kSeqGroup, kc - g0);
                for (int i = 0; i < gk && rc == CSLAM_OK; i++) {
                    const int st = base + g0 + i;
                    rc = cslam_ekf_predict(h, v[st], swa[st], Q, wb, dt);
                    if (rc == CSLAM_OK && use_heading) {
                        double* Ai = h->A + (size_t)i * h->lda;  // panel row i of this group
                        count_launch();
                        k_heading_gain<<<(n + 255) / 256, 256, 0, h->stream>>>(h->X[h->cur], h->X[h->cur ^ 1], h->R3,
                                                                               h->ld, n, phi[st], sigma * sigma, Ai);
                        h->cur ^= 1;
                        count_launch();
                        k_rows012_update<<<dim3((n + 255) / 256, 3), 256, 0, h->stream>>>(h->R3, h->ld, n, Ai, h->lda, 1,
                                                                                        kFltMin);
                        if (cudaGetLastError() != cudaSuccess) rc = CSLAM_ERR_CUDA;
                        if (rc == CSLAM_OK && h->sh.world > 1 && !h->diag_dirty) {
                            const int nf = (n - 3) / 2;
                            if (nf > 0) {
                                count_launch();
                                k_diag_follow<<<(nf + 255) / 256, 256, 0, h->stream>>>(h->D, h->dcap, nf, Ai, h->lda, 1, 1,
                                                                                       kFltMin, true);
                            }
                        }
                    }
                    if (rc == CSLAM_OK && trace_dev) {
                        cudaError_t e = cudaMemcpyAsync(trace_dev + 3 * st, h->X[h->cur], 3 * sizeof(double),
                                                        cudaMemcpyDeviceToDevice, h->stream);
                        if (e != cudaSuccess) rc = CSLAM_ERR_CUDA;
                    }
                }
                if (rc == CSLAM_OK && use_heading) rc = launch_heading_multi(h, gk);
            }
        }
    }
    if (rc == CSLAM_OK && pose_trace) {
        cudaError_t e = cudaMemcpyAsync(pose_trace, trace_dev, (size_t)3 * k * sizeof(double), cudaMemcpyDeviceToHost,
                                        h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) {
            set_last_error("cslam_ekf_control_steps: %s", cudaGetErrorString(e));
            rc = CSLAM_ERR_CUDA;
        }
    }
    return rc;
}

int cslam_ekf_gate(cslam_ekf_t* h, const double* Z, int m, const double R[4], double gate1, double gate2,
                   int32_t* jbest, uint8_t* is_new, double* nbest, double* outer) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_handle(h)) return rc;
    CSLAM_REQUIRE(m >= 0, CSLAM_ERR_BAD_ARG, "m < 0");
    CSLAM_REQUIRE(m == 0 || (Z && R && jbest), CSLAM_ERR_BAD_ARG, "null argument");
    const int nf = (h->n - 3) / 2;
    if (m > 0)
        if (int rc = refresh_diag_cache(h)) return rc;
    for (int base = 0; base < m; base += CSLAM_MAX_OBS) {
        const int mc = std::min(CSLAM_MAX_OBS, m - base);
        if (int rc = launch_gate(h->X[h->cur], h->P, h->R3, (h->sh.world > 1 || h->lz.on) ? h->D : nullptr, h->dcap, h->ld, nf,
                                 Z + 2 * base, mc, R, gate1, gate2, h->gate.part_nd, h->gate.part_out,
                                 h->gate.part_j, h->ticket + 1, h->gate.d_jbest, h->gate.d_nbest, h->gate.d_outer,
                                 nullptr, h->stream))
            return rc;
        char* pin = static_cast<char*>(h->pinned);
        int* pj = reinterpret_cast<int*>(pin);
        double* pn = reinterpret_cast<double*>(pin + 1024);
        double* po = reinterpret_cast<double*>(pin + 2048);
        CSLAM_CUDA(cudaMemcpyAsync(pj, h->gate.d_jbest, mc * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CSLAM_CUDA(cudaMemcpyAsync(pn, h->gate.d_nbest, mc * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CSLAM_CUDA(cudaMemcpyAsync(po, h->gate.d_outer, mc * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CSLAM_CUDA(cudaStreamSynchronize(h->stream));
        for (int i = 0; i < mc; i++) {
            jbest[base + i] = pj[i];
            if (is_new) is_new[base + i] = (pj[i] == 0 && po[i] > gate2) ? 1 : 0;  // EKF.cpp:287-295
            if (nbest) nbest[base + i] = pn[i];
            if (outer) outer[base + i] = po[i];
        }
    }
    return CSLAM_OK;
}

int cslam_ekf_update(cslam_ekf_t* h, const double* Z, const int32_t* idf, int m, const double R[4], int batch) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_handle(h)) return rc;
    CSLAM_REQUIRE(m >= 0, CSLAM_ERR_BAD_ARG, "m < 0");
    if (m == 0) return CSLAM_OK;  // test/main.cpp:188 calls update with an empty ZF
    CSLAM_REQUIRE(Z && idf && R, CSLAM_ERR_BAD_ARG, "null argument");
    const int n = h->n, nf = (n - 3) / 2;
    const bool sharded = h->sh.world > 1;
    for (int i = 0; i < m; i++)
        CSLAM_REQUIRE(idf[i] >= 1 && idf[i] <= nf, CSLAM_ERR_BAD_ARG, "idf out of range (1-based map slots)");
    if (!batch) return h->lz.on ? lazy_sequential(h, Z, idf, nullptr, m, R) : sequential_updates(h, Z, idf, nullptr, m, R);
    if (m > CSLAM_MAX_BATCH_OBS) {
        // One joint update handles rank <= 64.  More observations are applied as successive joint updates of up to
        // 32 each (every chunk linearised at the state the previous chunk produced) — documented in cslam.h; the
        // reference's batchUpdate has no such limit but its driver never sees more than 7 landmarks at once.
        for (int b = 0; b < m; b += CSLAM_MAX_BATCH_OBS)
            if (int rc = cslam_ekf_update(h, Z + 2 * b, idf + b, std::min(CSLAM_MAX_BATCH_OBS, m - b), R, 1)) return rc;
        return CSLAM_OK;
    }
    if (int rc = lazy_flush_all(h)) return rc;  // the joint update reads its 2m columns from the up-to-date array
    ObsPack ob;
    memset(&ob, 0, sizeof(ob));
    memcpy(ob.z, Z, sizeof(double) * 2 * m);
    memcpy(ob.idf, idf, sizeof(int) * m);
    ob.m = m;
    memcpy(ob.R, R, sizeof(double) * 4);
    const int r = 2 * m;
    count_launch();
    k_batch_prep<<<1, CSLAM_MAX_BATCH_OBS, 0, h->stream>>>(h->X[h->cur], ob, h->small);
    if (sharded || h->lz.on) {
        ColList cl;
        cl.n = r;
        for (int k = 0; k < m; k++) {
            cl.c[2 * k] = 3 + 2 * (idf[k] - 1);
            cl.c[2 * k + 1] = cl.c[2 * k] + 1;
        }
        const double* snap = h->colbuf;
        if (h->lz.on) {
            if (int rc = lazy_snapshot(h, cl, nullptr, &snap)) return rc;
        } else {
            if (int rc = exchange_columns(h, cl)) return rc;
        }
        count_launch();
        k_batch_pht<true><<<dim3((n + 127) / 128, m), 128, 0, h->stream>>>(h->P, h->R3, snap, h->ld, n, m,
                                                                            h->small, h->PHT, h->lda);
    } else {
        count_launch();
        k_batch_pht<false><<<dim3((n + 127) / 128, m), 128, 0, h->stream>>>(h->P, h->R3, nullptr, h->ld, n, m,
                                                                             h->small, h->PHT, h->lda);
    }
    const int chol_smem = 2 * kMaxRank * (kMaxRank + 1) * (int)sizeof(double);
    if (!h->attr_chol) {  // once per handle (device), not per call
        CSLAM_CUDA(cudaFuncSetAttribute(k_batch_chol, cudaFuncAttributeMaxDynamicSharedMemorySize, chol_smem));
        h->attr_chol = true;
    }
    count_launch();
    k_batch_chol<<<1, 256, chol_smem, h->stream>>>(h->PHT, h->lda, m, R[0], R[1], R[2], R[3], h->flags, h->small,
                                                   h->status);
    count_launch();
    k_batch_w1<<<dim3((n + 127) / 128, (r + 7) / 8), 128, 0, h->stream>>>(h->PHT, h->lda, n, r, h->small,
                                                                           h->X[h->cur], h->X[h->cur ^ 1], h->A);
    CSLAM_CUDA(cudaGetLastError());
    h->cur ^= 1;
    return launch_cov_update_rank(h, r);
}

int cslam_ekf_observe_step(cslam_ekf_t* h, const double* ZF, const int32_t* idf, int mf, const double* ZN, int mn,
                           const double R[4], int batch) {
    if (int rc = check_handle(h)) return rc;
    CSLAM_REQUIRE(mf >= 0 && mn >= 0, CSLAM_ERR_BAD_ARG, "negative count");
    CSLAM_REQUIRE(R && (mf == 0 || (ZF && idf)) && (mn == 0 || ZN), CSLAM_ERR_BAD_ARG, "null argument");
    const int n = h->n, nf = (n - 3) / 2;
    const bool fused = batch && !h->lz.on && h->sh.world == 1 && n + 2 * mn <= kSmallN && mf <= CSLAM_MAX_BATCH_OBS &&
                       mn <= CSLAM_MAX_BATCH_OBS && (mf > 0 || mn > 0);
    if (!fused) {  // big / sharded maps and sequential updates: the two calls of test/main.cpp:188-189
        if (int rc = cslam_ekf_update(h, ZF, idf, mf, R, batch)) return rc;
        return cslam_ekf_augment(h, ZN, mn, R);
    }
    for (int i = 0; i < mf; i++)
        CSLAM_REQUIRE(idf[i] >= 1 && idf[i] <= nf, CSLAM_ERR_BAD_ARG, "idf out of range (1-based map slots)");
    CSLAM_REQUIRE(n + 2 * mn <= h->n_cap, CSLAM_ERR_CAPACITY, "landmark capacity exceeded");
    StepPack sp;
    memset(&sp, 0, sizeof(sp));
    if (mf > 0) {
        memcpy(sp.ob.z, ZF, sizeof(double) * 2 * mf);
        memcpy(sp.ob.idf, idf, sizeof(int) * mf);
    }
    sp.ob.m = mf;
    memcpy(sp.ob.R, R, sizeof(double) * 4);
    if (mn > 0) memcpy(sp.zn, ZN, sizeof(double) * 2 * mn);
    sp.mn = mn;
    const int chol_smem = 2 * kMaxRank * (kMaxRank + 1) * (int)sizeof(double);
    if (!h->attr_step) {
        CSLAM_CUDA(cudaFuncSetAttribute(k_observe_step_small, cudaFuncAttributeMaxDynamicSharedMemorySize, chol_smem));
        h->attr_step = true;
    }
    count_launch();
    k_observe_step_small<<<1, 1024, chol_smem, h->stream>>>(h->X[h->cur], h->X[h->cur ^ 1], h->P, h->ld, n, sp, h->flags,
                                                            h->small, h->PHT, h->A, h->lda, h->status);
    CSLAM_CUDA(cudaGetLastError());
    if (mf > 0) h->cur ^= 1;
    h->n += 2 * mn;
    h->diag_dirty = true;
    return CSLAM_OK;
}

int cslam_ekf_scan(cslam_ekf_t* h, const double* Z, int m, const double R[4], double gate1, double gate2,
                   int32_t* jbest, uint8_t* is_new) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_handle(h)) return rc;
    CSLAM_REQUIRE(m >= 0 && m <= CSLAM_MAX_OBS, CSLAM_ERR_BAD_ARG, "m out of range (0..CSLAM_MAX_OBS)");
    if (m == 0) return CSLAM_OK;
    CSLAM_REQUIRE(Z && R, CSLAM_ERR_BAD_ARG, "null argument");
    const int n = h->n, nf = (n - 3) / 2;
    if (int rc = refresh_diag_cache(h)) return rc;
    // dataAssociate at the pre-update state for the whole scan (test/main.cpp:193), indices stay in d_jbest.
    // Deferred-pass handles: the gate kernel stops at its per-block candidates and the snapshot kernel of the
    // first update group merges them (gate_parts.cuh) — no ticket / last-block stage on the per-scan chain.
    const bool merge_later = h->lz.on && h->lz.fused_gains && h->lz.defer_gate_merge && nf > 0;
    int gate_blocks = 0;
    if (int rc = launch_gate(h->X[h->cur], h->P, h->R3, (h->sh.world > 1 || h->lz.on) ? h->D : nullptr, h->dcap, h->ld, nf, Z, m, R,
                             gate1, gate2,
                             h->gate.part_nd, h->gate.part_out, h->gate.part_j, h->ticket + 1, h->gate.d_jbest,
                             h->gate.d_nbest, h->gate.d_outer, h->assoc_count, h->stream, merge_later ? 0 : 1, &gate_blocks))
        return rc;
    auto queue_readback = [&]() -> int {  // optional read-back, queued behind the association only: overlaps the updates
        if (!(jbest || is_new)) return CSLAM_OK;
        char* pin = static_cast<char*>(h->pinned);
        CSLAM_CUDA(cudaMemcpyAsync(pin, h->gate.d_jbest, m * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CSLAM_CUDA(cudaMemcpyAsync(pin + 2048, h->gate.d_outer, m * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CSLAM_CUDA(cudaEventRecord(h->scan_ev, h->stream));
        return CSLAM_OK;
    };
    if (!merge_later)
        if (int rc = queue_readback()) return rc;
    // singleUpdate (EKF.cpp:457-479): re-linearised per observation, in observation order
    if (nf > 0) {
        if (h->lz.on) {
            GateParts gp;
            if (merge_later) {
                gp.nd = h->gate.part_nd;
                gp.out = h->gate.part_out;
                gp.j = h->gate.part_j;
                gp.nblocks = gate_blocks;
                gp.m = m;
                gp.jbest = h->gate.d_jbest;
                gp.nbest = h->gate.d_nbest;
                gp.outer = h->gate.d_outer;
                gp.assoc_count = h->assoc_count;
            }
            std::function<int()> rb = queue_readback;
            if (int rc = lazy_sequential(h, Z, nullptr, h->gate.d_jbest, m, R, merge_later ? &gp : nullptr,
                                         merge_later ? &rb : nullptr))
                return rc;
        } else if (int rc = sequential_updates(h, Z, nullptr, h->gate.d_jbest, m, R)) {
            return rc;
        }
    }
    if (jbest || is_new) {
        CSLAM_CUDA(cudaEventSynchronize(h->scan_ev));
        const int* pj = reinterpret_cast<const int*>(h->pinned);
        const double* po = reinterpret_cast<const double*>(static_cast<char*>(h->pinned) + 2048);
        for (int i = 0; i < m; i++) {
            if (jbest) jbest[i] = pj[i];
            if (is_new) is_new[i] = (pj[i] == 0 && po[i] > gate2) ? 1 : 0;  // EKF.cpp:287-295
        }
    }
    return CSLAM_OK;
}

int cslam_ekf_scan_associations(cslam_ekf_t* h, unsigned long long* total) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_handle(h)) return rc;
    CSLAM_REQUIRE(total != nullptr, CSLAM_ERR_BAD_ARG, "total is null");
    CSLAM_CUDA(cudaMemcpyAsync(h->pinned, h->assoc_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                               h->stream));
    CSLAM_CUDA(cudaStreamSynchronize(h->stream));
    *total = *static_cast<unsigned long long*>(h->pinned);
    return CSLAM_OK;
}

int cslam_ekf_augment(cslam_ekf_t* h, const double* Z, int m, const double R[4]) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_handle(h)) return rc;
    CSLAM_REQUIRE(m >= 0, CSLAM_ERR_BAD_ARG, "m < 0");
    if (m == 0) return CSLAM_OK;
    CSLAM_REQUIRE(Z && R, CSLAM_ERR_BAD_ARG, "null argument");
    CSLAM_REQUIRE(h->n + 2 * m <= h->n_cap, CSLAM_ERR_CAPACITY, "landmark capacity exceeded");
    // lazy handle: the new columns are written into the up-to-date array (pending panel rows end at the old n)
    if (int rc = lazy_flush_all(h)) return rc;
    const bool dcache = h->lz.on && !h->diag_dirty;  // enter the new 2x2 blocks into the cache directly
    for (int i = 0; i < m; i++) {
        const int len = h->n;
        count_launch();
        k_augment<<<(len + 255) / 256, 256, 0, h->stream>>>(h->X[h->cur], lazy_P(h), h->R3, h->ld, len, Z[2 * i],
                                                            Z[2 * i + 1], R[0], R[1], R[2], R[3], h->sh,
                                                            dcache ? h->D : nullptr, h->dcap);
        CSLAM_CUDA(cudaGetLastError());
        h->n += 2;
    }
    if (!dcache) h->diag_dirty = true;
    return CSLAM_OK;
}

int cslam_ekf_get_state(cslam_ekf_t* h, double* X, int max_n) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_handle(h)) return rc;
    CSLAM_REQUIRE(X != nullptr && max_n >= h->n, CSLAM_ERR_BAD_ARG, "buffer too small");
    CSLAM_CUDA(cudaMemcpyAsync(h->pinned, h->X[h->cur], h->n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CSLAM_CUDA(cudaStreamSynchronize(h->stream));
    memcpy(X, h->pinned, h->n * sizeof(double));
    return CSLAM_OK;
}

int cslam_ekf_get_cov_block(cslam_ekf_t* h, int r0, int c0, int nr, int nc, double* out) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_handle(h)) return rc;
    CSLAM_REQUIRE(out && r0 >= 0 && c0 >= 0 && nr >= 0 && nc >= 0 && r0 + nr <= h->n && c0 + nc <= h->n,
                  CSLAM_ERR_BAD_ARG, "block out of range");
    if (nr == 0 || nc == 0) return CSLAM_OK;
    if (int rc = lazy_flush_all(h)) return rc;
    const size_t cnt = (size_t)nr * nc;
    if (int rc = ensure_scratch(h, &h->acc_dev, &h->acc_cap, cnt)) return rc;
    double* tmp = h->acc_dev;
    count_launch();
    k_gather_block<<<(unsigned)((cnt + 255) / 256), 256, 0, h->stream>>>(lazy_P(h), h->R3, h->ld, r0, c0, nr, nc, tmp,
                                                                         h->sh);
    CSLAM_CUDA(cudaGetLastError());
    if (h->sh.world > 1)
        if (int rc = allreduce_sum(h, tmp, cnt)) return rc;  // collective: all ranks call
    CSLAM_CUDA(cudaMemcpyAsync(out, tmp, cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CSLAM_CUDA(cudaStreamSynchronize(h->stream));
    return CSLAM_OK;
}

// Arbitrary principal sub-matrix: out[a * k + b] = P(idx[a], idx[b]) — the joint marginal of a set of state
// entries (e.g. the pose and a handful of landmarks) without reading P back.
__global__ void k_gather_sub(const double* __restrict__ P, const double* __restrict__ R3, size_t ld,
                             const int* __restrict__ idx, int k, double* __restrict__ out, Shard sh) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)k * k) return;
    const int i = idx[t / k], j = idx[t % k];
    const int lo = i <= j ? i : j, hi = i <= j ? j : i;
    if (lo < 3)
        out[t] = sh.rank == 0 ? R3[(size_t)lo * ld + hi] : 0.0;
    else
        out[t] = shard_owns(sh, lo) ? P[shard_lrow(sh, lo) * ld + hi] : 0.0;
}

int cslam_ekf_get_cov_gather(cslam_ekf_t* h, const int32_t* idx, int k, double* out) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_handle(h)) return rc;
    CSLAM_REQUIRE(k >= 0 && k <= 8192, CSLAM_ERR_BAD_ARG, "k out of range (0..8192)");
    if (k == 0) return CSLAM_OK;
    CSLAM_REQUIRE(idx && out, CSLAM_ERR_BAD_ARG, "null argument");
    for (int a = 0; a < k; a++) CSLAM_REQUIRE(idx[a] >= 0 && idx[a] < h->n, CSLAM_ERR_BAD_ARG, "state index out of range");
    if (int rc = lazy_flush_all(h)) return rc;
    const size_t cnt = (size_t)k * k;
    // scratch: k*k doubles of output followed by the k indices
    if (int rc = ensure_scratch(h, &h->acc_dev, &h->acc_cap, cnt + (size_t)(k + 1) / 2)) return rc;
    int* didx = reinterpret_cast<int*>(h->acc_dev + cnt);
    CSLAM_CUDA(cudaMemcpyAsync(didx, idx, (size_t)k * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    count_launch();
    k_gather_sub<<<(unsigned)((cnt + 255) / 256), 256, 0, h->stream>>>(lazy_P(h), h->R3, h->ld, didx, k, h->acc_dev, h->sh);
    CSLAM_CUDA(cudaGetLastError());
    if (h->sh.world > 1)
        if (int rc = allreduce_sum(h, h->acc_dev, cnt)) return rc;  // collective: all ranks call
    CSLAM_CUDA(cudaMemcpyAsync(out, h->acc_dev, cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CSLAM_CUDA(cudaStreamSynchronize(h->stream));
    return CSLAM_OK;
}

// Landmark marginals for visualisation (the README's uncertainty ellipses) without pulling P:
// per landmark j0 <= j < j0 + count the 2x2 diagonal block packed as (P_ff, P_f,f+1, P_f+1,f+1).
__global__ void __launch_bounds__(256) k_landmark_covs(const double* __restrict__ P, size_t ld, int j0, int count,
                                                       double* __restrict__ out, Shard sh) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int f = 3 + 2 * (j0 + t);
    double a = 0.0, b = 0.0, c = 0.0;
    if (shard_owns(sh, f)) {
        a = P[shard_lrow(sh, f) * ld + f];
        b = P[shard_lrow(sh, f) * ld + f + 1];
    }
    if (shard_owns(sh, f + 1)) c = P[shard_lrow(sh, f + 1) * ld + f + 1];
    out[3 * (size_t)t] = a;
    out[3 * (size_t)t + 1] = b;
    out[3 * (size_t)t + 2] = c;
}

int cslam_ekf_get_landmark_covs(cslam_ekf_t* h, int first_landmark, int count, double* out) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_handle(h)) return rc;
    const int nf = (h->n - 3) / 2;
    CSLAM_REQUIRE(first_landmark >= 1 && count >= 0 && first_landmark - 1 + count <= nf, CSLAM_ERR_BAD_ARG,
                  "landmark range out of bounds (1-based ids)");
    if (count == 0) return CSLAM_OK;
    CSLAM_REQUIRE(out != nullptr, CSLAM_ERR_BAD_ARG, "out is null");
    if (int rc = lazy_flush_all(h)) return rc;
    const size_t cnt = 3 * (size_t)count;
    if (int rc = ensure_scratch(h, &h->acc_dev, &h->acc_cap, cnt)) return rc;
    double* tmp = h->acc_dev;
    count_launch();
    k_landmark_covs<<<(count + 255) / 256, 256, 0, h->stream>>>(lazy_P(h), h->ld, first_landmark - 1, count, tmp, h->sh);
    CSLAM_CUDA(cudaGetLastError());
    if (h->sh.world > 1)
        if (int rc = allreduce_sum(h, tmp, cnt)) return rc;  // collective: all ranks call
    CSLAM_CUDA(cudaMemcpyAsync(out, tmp, cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CSLAM_CUDA(cudaStreamSynchronize(h->stream));
    return CSLAM_OK;
}

// Checkpoint (the reference has none: X, P and mTABLE live in the driver).  One file per rank:
//   header {magic, version, flags, n, rank, world} + X[n] + rows 0..2 of P (3 x n, replicated on every rank)
//   + every covariance row i >= 3 this rank stores, ascending, from the diagonal on (n - i doubles each)
// — the device layout minus padding, 4 n (n+1) bytes of covariance in total over the ranks.  A sharded handle
// writes / reads  <path>.r<rank>of<world>  (every rank calls with the same path), a single-GPU handle <path>.
namespace {
struct CkptHeader {
    char magic[8];
    uint32_t version;
    uint32_t flags;
    int32_t n;
    int32_t rank;
    int32_t world;
    int32_t reserved;
};
const char kCkptMagic[8] = {'C', 'S', 'L', 'A', 'M', 'E', 'K', 'F'};
std::string ckpt_path(const cslam_ekf* h, const char* path) {
    std::string p(path);
    if (h->sh.world > 1) p += ".r" + std::to_string(h->sh.rank) + "of" + std::to_string(h->sh.world);
    return p;
}
long long ckpt_bytes(const cslam_ekf* h, int n) {
    long long rows = 0;
    for (int i = 3; i < n; i++)
        if (shard_owns(h->sh, i)) rows += n - i;
    return (long long)sizeof(CkptHeader) + 8LL * n + 8LL * 3 * n + 8LL * rows;
}
}  // namespace

int cslam_ekf_save(cslam_ekf_t* h, const char* path) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_handle(h)) return rc;
    CSLAM_REQUIRE(path != nullptr, CSLAM_ERR_BAD_ARG, "path is null");
    if (int rc = lazy_flush_all(h)) return rc;
    CSLAM_CUDA(cudaStreamSynchronize(h->stream));
    const double* Psrc = lazy_P(h);
    const std::string file = ckpt_path(h, path);
    FILE* f = fopen(file.c_str(), "wb");
    CSLAM_REQUIRE(f != nullptr, CSLAM_ERR_BAD_ARG, "cannot open checkpoint file for writing");
    const int n = h->n;
    CkptHeader hd;
    memset(&hd, 0, sizeof(hd));
    memcpy(hd.magic, kCkptMagic, 8);
    hd.version = 2;
    hd.flags = h->flags;
    hd.n = n;
    hd.rank = h->sh.rank;
    hd.world = h->sh.world;
    bool ok = fwrite(&hd, sizeof(hd), 1, f) == 1;
    std::vector<double> host((size_t)std::max(3 * n, 1));
    cudaError_t e = cudaMemcpy(host.data(), h->X[h->cur], n * sizeof(double), cudaMemcpyDeviceToHost);
    ok = ok && e == cudaSuccess && fwrite(host.data(), sizeof(double), n, f) == (size_t)n;
    if (ok) {  // rows 0..2: the always-current panel (aliases P on an eager single-GPU handle)
        e = cudaMemcpy2D(host.data(), (size_t)n * sizeof(double), h->R3, h->ld * sizeof(double), (size_t)n * sizeof(double), 3,
                         cudaMemcpyDeviceToHost);
        ok = e == cudaSuccess && fwrite(host.data(), sizeof(double), (size_t)3 * n, f) == (size_t)3 * n;
    }
    // owned rows in slabs of 128 (a shard block is contiguous in local storage), upper part written per row
    std::vector<double> rows((size_t)kShardRows * n);
    for (int i0 = 0; ok && i0 < n; i0 += kShardRows) {
        if (!shard_owns(h->sh, i0)) continue;
        const int nr = std::min(kShardRows, n - i0);
        e = cudaMemcpy2D(rows.data(), (size_t)n * sizeof(double), Psrc + shard_lrow(h->sh, i0) * h->ld, h->ld * sizeof(double),
                         (size_t)n * sizeof(double), nr, cudaMemcpyDeviceToHost);
        ok = e == cudaSuccess;
        for (int r = 0; ok && r < nr; r++) {
            const int i = i0 + r;
            if (i < 3) continue;
            ok = fwrite(rows.data() + (size_t)r * n + i, sizeof(double), n - i, f) == (size_t)(n - i);
        }
    }
    ok = (fclose(f) == 0) && ok;
    if (e != cudaSuccess) {
        set_last_error("cslam_ekf_save: %s", cudaGetErrorString(e));
        return CSLAM_ERR_CUDA;
    }
    CSLAM_REQUIRE(ok, CSLAM_ERR_BAD_ARG, "short write to the checkpoint file");
    return CSLAM_OK;
}

int cslam_ekf_load(cslam_ekf_t* h, const char* path) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_handle(h)) return rc;
    CSLAM_REQUIRE(path != nullptr, CSLAM_ERR_BAD_ARG, "path is null");
    const std::string file = ckpt_path(h, path);
    FILE* f = fopen(file.c_str(), "rb");
    CSLAM_REQUIRE(f != nullptr, CSLAM_ERR_BAD_ARG, "cannot open checkpoint file");
    CkptHeader hd;
    bool ok = fread(&hd, sizeof(hd), 1, f) == 1 && memcmp(hd.magic, kCkptMagic, 8) == 0 && hd.version == 2;
    if (!ok || hd.n < 3 || hd.n > h->n_cap || (hd.n - 3) % 2 != 0) {
        fclose(f);
        set_last_error("cslam_ekf_load: not a checkpoint of this library, or it exceeds the handle's capacity");
        return CSLAM_ERR_BAD_ARG;
    }
    // validate BEFORE touching device state: quirk mode, sharding and exact length (a truncated file must not
    // leave a half-overwritten handle behind)
    if (hd.flags != h->flags || hd.rank != h->sh.rank || hd.world != h->sh.world) {
        fclose(f);
        set_last_error("cslam_ekf_load: checkpoint written with flags 0x%x as rank %d of %d, handle has flags 0x%x, rank %d of %d",
                       hd.flags, hd.rank, hd.world, h->flags, h->sh.rank, h->sh.world);
        return CSLAM_ERR_BAD_ARG;
    }
    {
        const long long want = ckpt_bytes(h, hd.n);
        long long have = -1;
        if (fseek(f, 0, SEEK_END) == 0) have = ftell(f);
        if (have != want || fseek(f, (long)sizeof(hd), SEEK_SET) != 0) {
            fclose(f);
            set_last_error("cslam_ekf_load: checkpoint length %lld, expected %lld (truncated or corrupt)", have, want);
            return CSLAM_ERR_BAD_ARG;
        }
    }
    const int n = hd.n;
    CSLAM_CUDA(cudaStreamSynchronize(h->stream));
    if (h->lz.on) {  // the state is replaced: drop pending terms, array 0 becomes current
        LazyState& L = h->lz;
        CSLAM_CUDA(cudaStreamSynchronize(L.pass_stream));
        L.stable = L.newest = 0;
        L.np = L.infl_rows = 0;
        L.eps_mask = L.infl_eps_mask = 0;
        L.pass_pending_wait = false;
        L.stable_busy = false;
    }
    std::vector<double> x((size_t)std::max(3 * n, 1));
    ok = fread(x.data(), sizeof(double), n, f) == (size_t)n;
    cudaError_t e = cudaSuccess;
    if (ok) e = cudaMemcpy(h->X[h->cur], x.data(), n * sizeof(double), cudaMemcpyHostToDevice);
    if (ok && e == cudaSuccess) {
        ok = fread(x.data(), sizeof(double), (size_t)3 * n, f) == (size_t)3 * n;
        if (ok)
            e = cudaMemcpy2D(h->R3, h->ld * sizeof(double), x.data(), (size_t)n * sizeof(double), (size_t)n * sizeof(double), 3,
                             cudaMemcpyHostToDevice);
    }
    std::vector<double> rows((size_t)kShardRows * n, 0.0);
    for (int i0 = 0; ok && e == cudaSuccess && i0 < n; i0 += kShardRows) {
        if (!shard_owns(h->sh, i0)) continue;
        const int nr = std::min(kShardRows, n - i0);
        for (int r = 0; ok && r < nr; r++) {
            const int i = i0 + r;
            for (int j = 0; j < n; j++) rows[(size_t)r * n + j] = 0.0;  // below the diagonal / rows 0..2: unauthoritative
            if (i < 3) {
                if (h->R3 == h->P)  // eager single-GPU handle: rows 0..2 ARE the first rows of P
                    for (int j = i; j < n; j++) rows[(size_t)r * n + j] = x[(size_t)i * n + j];
                continue;
            }
            ok = fread(rows.data() + (size_t)r * n + i, sizeof(double), n - i, f) == (size_t)(n - i);
        }
        if (ok)
            e = cudaMemcpy2D(h->P + shard_lrow(h->sh, i0) * h->ld, h->ld * sizeof(double), rows.data(),
                             (size_t)n * sizeof(double), (size_t)n * sizeof(double), nr, cudaMemcpyHostToDevice);
    }
    fclose(f);
    if (e != cudaSuccess) {
        set_last_error("cslam_ekf_load: %s", cudaGetErrorString(e));
        return CSLAM_ERR_CUDA;
    }
    CSLAM_REQUIRE(ok, CSLAM_ERR_BAD_ARG, "truncated checkpoint file");
    h->n = n;
    h->diag_dirty = true;
    CSLAM_CUDA(cudaMemsetAsync(h->status, 0, sizeof(int), h->stream));
    return CSLAM_OK;
}

int cslam_ekf_reset(cslam_ekf_t* h, const double* X, int n, const double* P) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_handle(h)) return rc;
    CSLAM_REQUIRE(X && n >= 3 && n <= h->n_cap && ((n - 3) % 2 == 0), CSLAM_ERR_BAD_ARG, "bad state size");
    // everything on the handle's stream: the legacy default stream does not order against it
    CSLAM_CUDA(cudaMemsetAsync(h->X[0], 0, h->ld * sizeof(double), h->stream));
    CSLAM_CUDA(cudaMemsetAsync(h->X[1], 0, h->ld * sizeof(double), h->stream));
    CSLAM_CUDA(cudaMemcpyAsync(h->X[0], X, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    h->cur = 0;
    h->n = n;
    h->diag_dirty = true;
    if (h->lz.on) {  // drop whatever is pending or in flight: the state is replaced
        LazyState& L = h->lz;
        CSLAM_CUDA(cudaStreamSynchronize(L.pass_stream));
        L.stable = L.newest = 0;
        L.np = L.infl_rows = 0;
        L.eps_mask = L.infl_eps_mask = 0;
        L.pass_pending_wait = false;
        L.stable_busy = false;
    }
    CSLAM_CUDA(cudaMemsetAsync(h->P, 0, (size_t)h->local_rows_cap * h->ld * sizeof(double), h->stream));
    if (h->R3 != h->P) CSLAM_CUDA(cudaMemsetAsync(h->R3, 0, 3 * h->ld * sizeof(double), h->stream));
    CSLAM_CUDA(cudaMemsetAsync(h->status, 0, sizeof(int), h->stream));
    CSLAM_CUDA(cudaStreamSynchronize(h->stream));
    if (P) {
        const size_t cnt = (size_t)n * n;
        if (int rc = ensure_scratch(h, &h->acc_dev, &h->acc_cap, cnt)) return rc;
        CSLAM_CUDA(cudaMemcpyAsync(h->acc_dev, P, cnt * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        count_launch();
        k_scatter_upper<<<(unsigned)((cnt + 255) / 256), 256, 0, h->stream>>>(h->P, h->R3, h->ld, n, h->acc_dev, h->sh);
        CSLAM_CUDA(cudaGetLastError());
        CSLAM_CUDA(cudaStreamSynchronize(h->stream));
    }
    return CSLAM_OK;
}

int cslam_ekf_profile_begin(cslam_ekf_t* h, int max_launches) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_handle(h)) return rc;
    CSLAM_REQUIRE(max_launches > 0 && max_launches <= (1 << 20), CSLAM_ERR_BAD_ARG, "max_launches out of range");
    while ((int)h->prof_ev.size() < 2 * max_launches) {
        cudaEvent_t e;
        CSLAM_CUDA(cudaEventCreate(&e));
        h->prof_ev.push_back(e);
    }
    h->prof_used = 0;
    h->prof_bytes = 0.0;
    h->prof = true;
    return CSLAM_OK;
}

int cslam_ekf_profile_end(cslam_ekf_t* h, double* ms, int* launches, double* bytes) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_handle(h)) return rc;
    h->prof = false;
    CSLAM_CUDA(cudaStreamSynchronize(h->stream));
    if (h->lz.pass_stream) CSLAM_CUDA(cudaStreamSynchronize(h->lz.pass_stream));
    double total = 0.0;
    for (int i = 0; i + 1 < h->prof_used; i += 2) {
        float t = 0.f;
        CSLAM_CUDA(cudaEventElapsedTime(&t, h->prof_ev[i], h->prof_ev[i + 1]));
        total += t;
    }
    if (ms) *ms = total;
    if (launches) *launches = h->prof_used / 2;
    if (bytes) *bytes = h->prof_bytes;
    return CSLAM_OK;
}

int cslam_ekf_device_ptrs(cslam_ekf_t* h, void** dX, void** dP, size_t* ld) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_handle(h)) return rc;
    if (int rc = lazy_flush_all(h)) return rc;
    if (dX) *dX = h->X[h->cur];
    if (dP) *dP = lazy_P(h);  // lazy handles: rows 0..2 live in their own panel, not in this array
    if (ld) *ld = h->ld;
    return CSLAM_OK;
}

}  // extern "C"
