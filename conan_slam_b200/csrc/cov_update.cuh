// cov_update.cuh — the streaming covariance kernel  P <- P - W1 W1^T  (slam.h:260 and, with
// r = 1, the Joseph heading update slam.h:718): the north-star kernel, >99 % of an EKF update.
#pragma once
#include "common.cuh"

namespace cslam {

// cache-policy variants of the 128-bit accessors (HINT: 0 default, 1 streaming .cs, 2 L1::no_allocate)
template <int HINT>
__device__ __forceinline__ double2 cov_ld(const double* p) {
    if constexpr (HINT == 1) {
        return __ldcs(reinterpret_cast<const double2*>(p));
    } else if constexpr (HINT == 2) {
        double2 v;
        asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
        return v;
    } else {
        return ld128(p);
    }
}
template <int HINT>
__device__ __forceinline__ void cov_st(double* p, double2 v) {
    if constexpr (HINT == 1) {
        __stcs(reinterpret_cast<double2*>(p), v);
    } else {
        st128(p, v);
    }
}

// slam.h:260  P <- P - W1 W1^T over the UPPER TRIANGLE only, in place, FP64.
// One CTA per T x T tile of the triangle (tiles with tc >= tr); thread = one 16-byte column
// pair x T/RG rows, all loads of a batch issued before the first use so that every SM keeps
// tens of KB in flight; the panel rows of the tile are staged in shared memory (broadcast
// reads), the two panel columns a thread owns stay in registers.  Elements below the
// diagonal inside diagonal tiles are neither loaded nor stored.  diag_eps implements
// slam.h:719 (P += I * FLT_MIN) for the heading update.
// live (nullable): device flag of a fused scan, 0 = skip this update (EKF.cpp:287-295 found no match).
// Template knobs: R rank (1 or 2), T tile edge, BATCH loads in flight per thread, MINB minimum
// resident CTAs per SM (register cap), HINT cache policy of the P accesses.
template <int R, int T, int BATCH_ = 8, int MINB = 2, int HINT = 0>
__global__ void __launch_bounds__(256, MINB) k_cov_update(double* __restrict__ P, size_t ld, int n,
                                                          const double* __restrict__ A, size_t lda, int nt,
                                                          double diag_eps, Shard sh,
                                                          const int* __restrict__ live = nullptr) {
    constexpr int CP = T / 2;       // column pairs per tile
    constexpr int RG = 256 / CP;    // row groups
    constexpr int RPT = T / RG;     // rows per thread
    constexpr int BATCH = RPT > BATCH_ ? BATCH_ : RPT;
    __shared__ double sAr[R][T];
    // device-side association (cslam_ekf_scan): *live == 0 means "no landmark passed the gate for
    // this observation" and the whole update is a no-op; the flag is fetched together with the
    // panels so that its latency is not a serial prefix of every CTA
    const int live_v = live != nullptr ? *live : 1;

    int tr, tc;
    shard_tile(blockIdx.x, nt, sh, tr, tc);

    const int i0 = tr * T, j0 = tc * T;
    for (int idx = threadIdx.x; idx < R * T; idx += 256) {
        const int k = idx / T, ii = idx % T;
        sAr[k][ii] = (i0 + ii < n) ? A[(size_t)k * lda + i0 + ii] : 0.0;
    }
    const int cp = threadIdx.x % CP, rg = threadIdx.x / CP;
    const int j = j0 + 2 * cp;
    // a tile never straddles a 128-row shard block, so its rows are consecutive in local storage: one
    // division per CTA instead of one per row access (the streaming passes are close to issue-bound)
    double* __restrict__ const Pt = P + shard_lrow(sh, i0) * ld + j;
    double aj0[R], aj1[R];
#pragma unroll
    for (int k = 0; k < R; k++) {
        aj0[k] = (j < n) ? A[(size_t)k * lda + j] : 0.0;
        aj1[k] = (j + 1 < n) ? A[(size_t)k * lda + j + 1] : 0.0;
    }
    __syncthreads();
    if (j >= n || live_v == 0) return;
    const bool diag_tile = (tr == tc);
#pragma unroll 1
    for (int b0 = 0; b0 < RPT; b0 += BATCH) {
        double2 v[BATCH];
#pragma unroll
        for (int b = 0; b < BATCH; b++) {
            const int ii = rg + (b0 + b) * RG;
            const int i = i0 + ii;
            const bool act = (i < n) && (!diag_tile || j + 1 >= i);
            if (act) v[b] = cov_ld<HINT>(Pt + (size_t)(i - i0) * ld);
        }
#pragma unroll
        for (int b = 0; b < BATCH; b++) {
            const int ii = rg + (b0 + b) * RG;
            const int i = i0 + ii;
            const bool act = (i < n) && (!diag_tile || j + 1 >= i);
            if (act) {
                // dot product over the panel rows, first product taken as is (no 0 + x: one DADD
                // less per element; differs from a zero-initialised sum only in the sign of a zero)
                double s0 = sAr[0][ii] * aj0[0], s1 = sAr[0][ii] * aj1[0];
#pragma unroll
                for (int k = 1; k < R; k++) {
                    const double ai = sAr[k][ii];
                    s0 += ai * aj0[k];
                    s1 += ai * aj1[k];
                }
                double2 o = v[b];
                if (j >= i) o.x = o.x - s0;
                if (j + 1 < n) o.y = o.y - s1;
                if (j == i) o.x += diag_eps;
                if (j + 1 == i) o.y += diag_eps;
                cov_st<HINT>(Pt + (size_t)(i - i0) * ld, o);
            }
        }
    }
}

// The rank-2 term one sequential update subtracts from P(i, j):  a0_i a0_j + a1_i a1_j  — the
// operation order of k_cov_update<2> above (no FMA contraction in this translation unit), shared
// with the gain kernel's view of "P after the pending updates".
__device__ __forceinline__ double rank2_term(double a0i, double a1i, double a0j, double a1j) {
    double s = a0i * a0j;
    s += a1i * a1j;
    return s;
}

// M consecutive SEQUENTIAL rank-2 updates (EKF.cpp:457-479: one per observation, each gain
// re-linearised after the previous one) applied in ONE pass over the upper triangle:
//   P(i,j) <- ((P(i,j) - t_0) - t_1) ... - t_{M-1},   t_q = rank2_term of panel rows 2q, 2q+1
// which is, operation for operation, what M launches of k_cov_update<2> compute — the result is
// bit-identical, the covariance is read and written once instead of M times.  The gain kernel of
// observation q sees P "after updates 0..q-1" by applying the same terms to the few entries it reads
// (k_gain_single, kprev).  live: M device flags (nullable); a skipped observation has a zero panel.
template <int M, int T, int BATCH_ = 4, int MINB = 2, int HINT = 1>
__global__ void __launch_bounds__(256, MINB) k_cov_update_multi(double* __restrict__ P, size_t ld, int n,
                                                                const double* __restrict__ A, size_t lda, int nt,
                                                                Shard sh, const int* __restrict__ live) {
    constexpr int CP = T / 2, RG = 256 / CP, RPT = T / RG;
    constexpr int BATCH = RPT > BATCH_ ? BATCH_ : RPT;
    // Both panels of the tile live in shared memory, so a thread keeps only ONE update's four column
    // values in registers at a time (the register budget decides how many CTAs stream concurrently):
    //   sRow[i][q]  = (a0_i, a1_i) of update q   — read as one broadcast 16-byte load per (row, q)
    //   sCol0/1[q][cp] = (a0_j, a0_j+1) / (a1_j, a1_j+1) — one conflict-free 16-byte load each per (batch, q)
    __shared__ double2 sRow[T][M];
    __shared__ double2 sCol0[M][CP], sCol1[M][CP];
    int any_live = 1;
    if (live != nullptr) {
        any_live = 0;
#pragma unroll
        for (int q = 0; q < M; q++) any_live |= live[q];
    }
    int tr, tc;
    shard_tile(blockIdx.x, nt, sh, tr, tc);
    const int i0 = tr * T, j0 = tc * T;
    for (int idx = threadIdx.x; idx < M * T; idx += 256) {
        const int q = idx / T, ii = idx % T;
        const bool in = i0 + ii < n;
        sRow[ii][q] = make_double2(in ? A[(size_t)(2 * q) * lda + i0 + ii] : 0.0,
                                   in ? A[(size_t)(2 * q + 1) * lda + i0 + ii] : 0.0);
    }
    for (int idx = threadIdx.x; idx < M * CP; idx += 256) {
        const int q = idx / CP, c = idx % CP;
        const int jj = j0 + 2 * c;
        const double* a0 = A + (size_t)(2 * q) * lda;
        const double* a1 = a0 + lda;
        sCol0[q][c] = make_double2(jj < n ? a0[jj] : 0.0, jj + 1 < n ? a0[jj + 1] : 0.0);
        sCol1[q][c] = make_double2(jj < n ? a1[jj] : 0.0, jj + 1 < n ? a1[jj + 1] : 0.0);
    }
    const int cp = threadIdx.x % CP, rg = threadIdx.x / CP;
    const int j = j0 + 2 * cp;
    // a tile never straddles a 128-row shard block, so its rows are consecutive in local storage: one
    // division per CTA instead of one per row access (the streaming passes are close to issue-bound)
    double* __restrict__ const Pt = P + shard_lrow(sh, i0) * ld + j;
    __syncthreads();
    if (j >= n || any_live == 0) return;
    const bool diag_tile = (tr == tc);
    const bool y_in = j + 1 < n;
#pragma unroll 1
    for (int b0 = 0; b0 < RPT; b0 += BATCH) {
        double2 v[BATCH];
#pragma unroll
        for (int b = 0; b < BATCH; b++) {
            const int i = i0 + rg + (b0 + b) * RG;
            const bool act = (i < n) && (!diag_tile || j + 1 >= i);
            if (act) v[b] = cov_ld<HINT>(Pt + (size_t)(i - i0) * ld);
        }
        if (diag_tile) {  // only tiles on the diagonal need the j >= i mask
#pragma unroll
            for (int q = 0; q < M; q++) {
                const double2 c0 = sCol0[q][cp], c1 = sCol1[q][cp];
#pragma unroll
                for (int b = 0; b < BATCH; b++) {
                    const int ii = rg + (b0 + b) * RG;
                    const int i = i0 + ii;
                    const double2 r = sRow[ii][q];
                    const double s0 = rank2_term(r.x, r.y, c0.x, c1.x);
                    const double s1 = rank2_term(r.x, r.y, c0.y, c1.y);
                    if (j >= i) v[b].x = v[b].x - s0;
                    if (y_in) v[b].y = v[b].y - s1;
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < M; q++) {
                const double2 c0 = sCol0[q][cp], c1 = sCol1[q][cp];
#pragma unroll
                for (int b = 0; b < BATCH; b++) {
                    const double2 r = sRow[rg + (b0 + b) * RG][q];
                    v[b].x = v[b].x - rank2_term(r.x, r.y, c0.x, c1.x);
                    if (y_in) v[b].y = v[b].y - rank2_term(r.x, r.y, c0.y, c1.y);
                }
            }
        }
#pragma unroll
        for (int b = 0; b < BATCH; b++) {
            const int i = i0 + rg + (b0 + b) * RG;
            const bool act = (i < n) && (!diag_tile || j + 1 >= i);
            if (act) cov_st<HINT>(Pt + (size_t)(i - i0) * ld, v[b]);
        }
    }
}


// K consecutive heading updates (EKF.cpp:328-352 -> slam.h:700-725, each a rank-1 pass P -= a a^T with
// FLT_MIN added on the diagonal, slam.h:719) applied in ONE pass over rows >= row_min of the upper
// triangle:  P(i,j) <- ((P(i,j) - a0_i a0_j [+ eps]) - a1_i a1_j [+ eps]) ...  — the operations of K
// launches of k_cov_update<1>, bit-identical.  The predicts between the heading updates of consecutive
// control steps only rewrite rows 0..2 (EKF.cpp:439-443 on upper-triangle storage), which the caller
// keeps current eagerly (k_rows012_update) and excludes here with row_min = 3; every other row sees
// nothing but the K rank-1 terms, so their passes commute with the predicts and merge.
// Panel rows 0..K-1 of A are the K vectors a = p / sqrt(S).
template <int K, int T, int BATCH_ = 4, int MINB = 4, int HINT = 1>
__global__ void __launch_bounds__(256, MINB) k_cov_update_multi1(double* __restrict__ P, size_t ld, int n,
                                                                 const double* __restrict__ A, size_t lda, int nt,
                                                                 double diag_eps, int row_min, Shard sh) {
    constexpr int CP = T / 2, RG = 256 / CP, RPT = T / RG;
    constexpr int BATCH = RPT > BATCH_ ? BATCH_ : RPT;
    __shared__ double sRow[T][K];
    __shared__ double2 sCol[K][CP];
    int tr, tc;
    shard_tile(blockIdx.x, nt, sh, tr, tc);
    const int i0 = tr * T, j0 = tc * T;
    for (int idx = threadIdx.x; idx < K * T; idx += 256) {
        const int q = idx / T, ii = idx % T;
        sRow[ii][q] = (i0 + ii < n) ? A[(size_t)q * lda + i0 + ii] : 0.0;
    }
    for (int idx = threadIdx.x; idx < K * CP; idx += 256) {
        const int q = idx / CP, c = idx % CP;
        const int jj = j0 + 2 * c;
        const double* a = A + (size_t)q * lda;
        sCol[q][c] = make_double2(jj < n ? a[jj] : 0.0, jj + 1 < n ? a[jj + 1] : 0.0);
    }
    const int cp = threadIdx.x % CP, rg = threadIdx.x / CP;
    const int j = j0 + 2 * cp;
    // a tile never straddles a 128-row shard block, so its rows are consecutive in local storage: one
    // division per CTA instead of one per row access (the streaming passes are close to issue-bound)
    double* __restrict__ const Pt = P + shard_lrow(sh, i0) * ld + j;
    __syncthreads();
    if (j >= n) return;
    const bool diag_tile = (tr == tc);
    const bool y_in = j + 1 < n;
    if (!diag_tile && i0 >= row_min && i0 + T <= n && j0 + T <= n) {
        // interior tile (almost all of them): no masks, one pointer walking down the rows.  (Measured: helps
        // this rank-1 pass, 4.95 -> 4.54 ms per drive cycle; the same fast path made the rank-2 pass slower.)
        double* __restrict__ p = Pt + (size_t)rg * ld;
        const size_t step = (size_t)RG * ld;
#pragma unroll 1
        for (int b0 = 0; b0 < RPT; b0 += BATCH) {
            double2 v[BATCH];
#pragma unroll
            for (int b = 0; b < BATCH; b++) v[b] = cov_ld<HINT>(p + b * step);
#pragma unroll
            for (int q = 0; q < K; q++) {
                const double2 c = sCol[q][cp];
#pragma unroll
                for (int b = 0; b < BATCH; b++) {
                    const double r = sRow[rg + (b0 + b) * RG][q];
                    v[b].x = v[b].x - r * c.x;
                    v[b].y = v[b].y - r * c.y;
                }
            }
#pragma unroll
            for (int b = 0; b < BATCH; b++) cov_st<HINT>(p + b * step, v[b]);
            p += BATCH * step;
        }
        return;
    }
#pragma unroll 1
    for (int b0 = 0; b0 < RPT; b0 += BATCH) {
        double2 v[BATCH];
#pragma unroll
        for (int b = 0; b < BATCH; b++) {
            const int i = i0 + rg + (b0 + b) * RG;
            const bool act = (i < n) && (i >= row_min) && (!diag_tile || j + 1 >= i);
            if (act) v[b] = cov_ld<HINT>(Pt + (size_t)(i - i0) * ld);
        }
        if (diag_tile) {  // only tiles on the diagonal carry the per-term FLT_MIN and the j >= i mask
#pragma unroll
            for (int q = 0; q < K; q++) {
                const double2 c = sCol[q][cp];
#pragma unroll
                for (int b = 0; b < BATCH; b++) {
                    const int ii = rg + (b0 + b) * RG;
                    const int i = i0 + ii;
                    const double r = sRow[ii][q];
                    if (j >= i) v[b].x = v[b].x - r * c.x;
                    if (y_in) v[b].y = v[b].y - r * c.y;
                    if (j == i) v[b].x += diag_eps;
                    if (j + 1 == i) v[b].y += diag_eps;
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < K; q++) {
                const double2 c = sCol[q][cp];
#pragma unroll
                for (int b = 0; b < BATCH; b++) {
                    const double r = sRow[rg + (b0 + b) * RG][q];
                    v[b].x = v[b].x - r * c.x;
                    if (y_in) v[b].y = v[b].y - r * c.y;
                }
            }
        }
#pragma unroll
        for (int b = 0; b < BATCH; b++) {
            const int i = i0 + rg + (b0 + b) * RG;
            const bool act = (i < n) && (i >= row_min) && (!diag_tile || j + 1 >= i);
            if (act) cov_st<HINT>(Pt + (size_t)(i - i0) * ld, v[b]);
        }
    }
}

}  // namespace cslam
