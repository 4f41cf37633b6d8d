// ekf_dmma.cu — joint (batch) covariance update  P <- P - W1 W1^T  (slam.h:260) for a
// rank-r panel (r = 2m <= 64) on the FP64 tensor cores of sm_100a.
//
// FP64 is not a tcgen05.mma kind: Blackwell's FP64 tensor path is the warp-level
// mma.sync.aligned.m8n8k4.f64 (SASS: DMMA).  The update is a real rank-r contraction
// (8 flop/B at r = 64, above the FP64 ridge) only in this batched case; the sequential
// rank-1/2 updates stay on the streaming kernel in ekf.cu.
//
// Tiling: one CTA = 64 x 128 tile of the upper triangle (tiles with any j >= i), 8 warps,
// each warp a 32 x 32 sub-tile = 4 x 4 DMMA blocks.  The accumulator fragments ARE the
// covariance: they are loaded straight from P (C operand), the negated row panel is the A
// operand, the column panel the B operand, and the fragments are stored back — P is read
// once and written once.  Panels are staged in shared memory with a row stride = 4 (mod 16)
// doubles, which makes the fragment reads bank-conflict free.
#include "common.cuh"
#include "cov_update.cuh"

namespace cslam {

constexpr int DM_TM = 64, DM_TN = 128;
constexpr int DM_SR = DM_TM + 4;  // 68
constexpr int DM_SC = DM_TN + 4;  // 132

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256, 2) k_cov_update_dmma(double* __restrict__ P, size_t ld, int n,
                                                            const double* __restrict__ A, size_t lda, int r,
                                                            int rp, int nbc, Shard sh) {
    extern __shared__ double smem[];
    double* sR = smem;               // [rp][DM_SR]  negated row panel
    double* sC = smem + rp * DM_SR;  // [rp][DM_SC]  column panel

    // linear tile id -> (br, bc): every owned 128-row shard tile `tr` holds two 64-row tile rows
    // (2tr, 2tr+1) that both start at column tile bc = tr  (nbc - tr tiles each)
    const long long t = blockIdx.x;
    int tr, tcd;
    shard_tile(t >> 1, nbc, sh, tr, tcd);
    const long long first = 2 * shard_first_tile((tr - sh.rank) / sh.world, nbc, sh);
    const int rem = (int)(t - first), cnt = nbc - tr;
    const int br = rem < cnt ? 2 * tr : 2 * tr + 1;
    const int bc = rem < cnt ? tr + rem : tr + rem - cnt;
    const int i0 = br * DM_TM, j0 = bc * DM_TN;

    for (int idx = threadIdx.x; idx < rp * DM_TM; idx += 256) {
        const int k = idx / DM_TM, ii = idx % DM_TM;
        sR[k * DM_SR + ii] = (k < r && i0 + ii < n) ? -A[(size_t)k * lda + i0 + ii] : 0.0;
    }
    for (int idx = threadIdx.x; idx < rp * DM_TN; idx += 256) {
        const int k = idx / DM_TN, jj = idx % DM_TN;
        sC[k * DM_SC + jj] = (k < r && j0 + jj < n) ? A[(size_t)k * lda + j0 + jj] : 0.0;
    }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wr = warp >> 2, wc = warp & 3;
    const int lr = lane >> 2, lc = lane & 3;
    const int iw = i0 + wr * 32, jw = j0 + wc * 32;
    // a warp sub-tile entirely below the diagonal has nothing to do
    const bool warp_active = (jw + 31 >= iw) && (iw < n) && (jw < n);

    double acc[4][4][2];
    if (warp_active) {
#pragma unroll
        for (int bi = 0; bi < 4; bi++) {
            const int i = iw + bi * 8 + lr;
#pragma unroll
            for (int bj = 0; bj < 4; bj++) {
                const int j = jw + bj * 8 + 2 * lc;
                double2 v = make_double2(0.0, 0.0);
                if (i < n && j < n && j + 1 >= i) v = ld128(P + shard_lrow(sh, i) * ld + j);
                acc[bi][bj][0] = v.x;
                acc[bi][bj][1] = v.y;
            }
        }
    }
    __syncthreads();
    if (!warp_active) return;

    const double* pr = sR + lc * DM_SR + wr * 32 + lr;
    const double* pc = sC + lc * DM_SC + wc * 32 + lr;
    for (int k0 = 0; k0 < rp; k0 += 4) {
        double a[4], b[4];
#pragma unroll
        for (int x = 0; x < 4; x++) {
            a[x] = pr[k0 * DM_SR + x * 8];
            b[x] = pc[k0 * DM_SC + x * 8];
        }
#pragma unroll
        for (int bi = 0; bi < 4; bi++)
#pragma unroll
            for (int bj = 0; bj < 4; bj++) dmma884(acc[bi][bj][0], acc[bi][bj][1], a[bi], b[bj]);
    }
#pragma unroll
    for (int bi = 0; bi < 4; bi++) {
        const int i = iw + bi * 8 + lr;
#pragma unroll
        for (int bj = 0; bj < 4; bj++) {
            const int j = jw + bj * 8 + 2 * lc;
            // pairs straddling the diagonal (j + 1 == i) rewrite one unauthoritative lower element
            if (i < n && j < n && j + 1 >= i)
                st128(P + shard_lrow(sh, i) * ld + j, make_double2(acc[bi][bj][0], acc[bi][bj][1]));
        }
    }
}

int launch_cov_update_dmma(double* P, size_t ld, int n, const double* A, size_t lda, int r, Shard sh,
                           cudaStream_t stream) {
    const int rp = (r + 3) / 4 * 4;
    const int nbc = (n + DM_TN - 1) / DM_TN;
    const long long tiles = 2 * shard_tile_count(nbc, sh);  // a trailing half-empty tile row exits early
    if (tiles == 0) return CSLAM_OK;
    const size_t smem = (size_t)rp * (DM_SR + DM_SC) * sizeof(double);
    CSLAM_CUDA(cudaFuncSetAttribute(k_cov_update_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    64 * (DM_SR + DM_SC) * (int)sizeof(double)));
    count_launch();
    k_cov_update_dmma<<<(unsigned)tiles, 256, smem, stream>>>(P, ld, n, A, lda, r, rp, nbc, sh);
    CSLAM_CUDA(cudaGetLastError());
    return CSLAM_OK;
}

}  // namespace cslam
