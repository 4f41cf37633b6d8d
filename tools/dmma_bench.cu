// tools/dmma_bench.cu — development benchmark for the FP64 tensor-core joint update:
//   (1) DMMA.8x8x4 peak of the GPU (register-resident chains),
//   (2) the production kernel (ekf_dmma.cu) against the first-round kernel (kept here only as a
//       bit-exact cross-check and speed reference), several chunk lengths.
//   (3) ablation of the production kernel: dbg & 1 skips the covariance loads, dbg & 2 the stores.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I conan_slam_b200/csrc \
//        tools/dmma_bench.cu conan_slam_b200/lib/util.o -ldl -o tools/bin/dmma_bench
#define CSLAM_DMMA_ABLATION 1
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../conan_slam_b200/csrc/ekf_dmma.cu"

using namespace cslam;
namespace cslam { extern int g_dmma_dbg; }

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

namespace cslam {
constexpr int V1_TM = 64, V1_TN = 128;
constexpr int V1_SR = V1_TM + 4;  // 68
constexpr int V1_SC = V1_TN + 4;  // 132

__device__ __forceinline__ void dmma884_v1(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256, 2) k_v1(double* __restrict__ P, size_t ld, int n,
                                                            const double* __restrict__ A, size_t lda, int r,
                                                            int rp, int nbc, Shard sh) {
    extern __shared__ double smem[];
    double* sR = smem;               // [rp][V1_SR]  negated row panel
    double* sC = smem + rp * V1_SR;  // [rp][V1_SC]  column panel

    // linear tile id -> (br, bc): every owned 128-row shard tile `tr` holds two 64-row tile rows
    // (2tr, 2tr+1) that both start at column tile bc = tr  (nbc - tr tiles each)
    const long long t = blockIdx.x;
    int tr, tcd;
    shard_tile(t >> 1, nbc, sh, tr, tcd);
    const long long first = 2 * shard_first_tile((tr - sh.rank) / sh.world, nbc, sh);
    const int rem = (int)(t - first), cnt = nbc - tr;
    const int br = rem < cnt ? 2 * tr : 2 * tr + 1;
    const int bc = rem < cnt ? tr + rem : tr + rem - cnt;
    const int i0 = br * V1_TM, j0 = bc * V1_TN;

    for (int idx = threadIdx.x; idx < rp * V1_TM; idx += 256) {
        const int k = idx / V1_TM, ii = idx % V1_TM;
        sR[k * V1_SR + ii] = (k < r && i0 + ii < n) ? -A[(size_t)k * lda + i0 + ii] : 0.0;
    }
    for (int idx = threadIdx.x; idx < rp * V1_TN; idx += 256) {
        const int k = idx / V1_TN, jj = idx % V1_TN;
        sC[k * V1_SC + jj] = (k < r && j0 + jj < n) ? A[(size_t)k * lda + j0 + jj] : 0.0;
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

    const double* pr = sR + lc * V1_SR + wr * 32 + lr;
    const double* pc = sC + lc * V1_SC + wc * 32 + lr;
    for (int k0 = 0; k0 < rp; k0 += 4) {
        double a[4], b[4];
#pragma unroll
        for (int x = 0; x < 4; x++) {
            a[x] = pr[k0 * V1_SR + x * 8];
            b[x] = pc[k0 * V1_SC + x * 8];
        }
#pragma unroll
        for (int bi = 0; bi < 4; bi++)
#pragma unroll
            for (int bj = 0; bj < 4; bj++) dmma884_v1(acc[bi][bj][0], acc[bi][bj][1], a[bi], b[bj]);
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

}  // namespace cslam

__global__ void __launch_bounds__(512) k_peak(double* out, int iters) {
    double c[16][2];
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
#pragma unroll
    for (int i = 0; i < 16; i++) c[i][0] = c[i][1] = 0.0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

__global__ void k_fill(double* P, size_t ld, int n) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = ld * (size_t)n;
    for (size_t t = idx; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(t / ld), j = (int)(t % ld);
        P[t] = (double)((i * 131 + j * 31) % 1009) * 1e-3 + (i == j ? 100.0 : 0.0);
    }
}
__global__ void k_fillA(double* A, size_t lda, int n, int r) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (i < n && k < r) A[(size_t)k * lda + i] = (double)((i * 7 + k * 13) % 257) * 1e-4 - 0.0128;
}
__global__ void k_cmp(const double* P1, const double* P2, size_t ld, int n, unsigned long long* bad) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = ld * (size_t)n;
    unsigned long long c = 0;
    for (size_t t = idx; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(t / ld), j = (int)(t % ld);
        if (j >= i && j < n && P1[t] != P2[t]) c++;
    }
    if (c) atomicAdd(bad, c);
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 40003;
    const int reps = argc > 2 ? atoi(argv[2]) : 5;
    const int r = argc > 3 ? atoi(argv[3]) : 64;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    // ---- (1) DMMA peak
    {
        double* d; CK(cudaMalloc(&d, 8));
        const int iters = 4096;
        for (int warps : {2, 4, 8, 16}) {
            k_peak<<<148, warps * 32>>>(d, 16);
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            k_peak<<<148 * 4, warps * 32>>>(d, iters);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            CK(cudaEventElapsedTime(&ms, e0, e1));
            const double flop = 148.0 * 4 * warps * iters * 16 * 512.0;
            printf("DMMA.8x8x4 peak: %2d warps/CTA, 4 CTAs/SM-slot: %7.2f TFLOP/s\n", warps, flop / (ms * 1e-3) / 1e12);
        }
        cudaFree(d);
    }
    // ---- (2) kernels
    const size_t ld = ((size_t)n + 1 + 15) / 16 * 16;
    double *P1, *P2, *A, *panels;
    unsigned long long* bad;
    CK(cudaMalloc(&P1, ld * n * sizeof(double)));
    CK(cudaMalloc(&P2, ld * n * sizeof(double)));
    CK(cudaMalloc(&A, 64 * ld * sizeof(double)));
    CK(cudaMalloc(&panels, dmma_panel_doubles(n) * sizeof(double)));
    CK(cudaMalloc(&bad, 8));
    CK(cudaMemset(A, 0, 64 * ld * sizeof(double)));
    CK(cudaMemset(panels, 0, dmma_panel_doubles(n) * sizeof(double)));
    k_fill<<<148 * 8, 256>>>(P1, ld, n);
    k_fill<<<148 * 8, 256>>>(P2, ld, n);
    k_fillA<<<dim3((n + 255) / 256, r), 256>>>(A, ld, n, r);
    CK(cudaDeviceSynchronize());
    const Shard sh{0, 1};
    const double flop = (double)r * n * ((double)n + 1.0), gb = 8.0 * n * ((double)n + 1.0) / 1e9;
    auto run_v1 = [&](double* P) {
        const int rp = (r + 3) / 4 * 4;
        const int nbc = (n + V1_TN - 1) / V1_TN;
        const long long tiles = 2 * shard_tile_count(nbc, sh);
        const size_t smem = (size_t)rp * (V1_SR + V1_SC) * sizeof(double);
        CK(cudaFuncSetAttribute(k_v1, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * (V1_SR + V1_SC) * 8));
        k_v1<<<(unsigned)tiles, 256, smem>>>(P, ld, n, A, ld, r, rp, nbc, sh);
    };
    // correctness: one application of each kernel on identical inputs must agree bit for bit
    run_v1(P1);
    if (launch_cov_update_dmma(P2, ld, n, A, ld, r, sh, panels, n, 32, 0)) { printf("launch failed: %s\n", cslam_last_error()); return 1; }
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(bad, 0, 8));
    k_cmp<<<148 * 8, 256>>>(P1, P2, ld, n, bad);
    unsigned long long hb = 0;
    CK(cudaMemcpy(&hb, bad, 8, cudaMemcpyDeviceToHost));
    printf("n=%d r=%d: v1 vs production kernel mismatching upper-triangle elements: %llu\n", n, r, hb);
    // timing
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; i++) run_v1(P1);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("v1  (round-1 kernel)         : %8.3f ms  %6.2f TFLOP/s  %7.1f GB/s\n", ms / reps, flop / (ms / reps * 1e-3) / 1e12, gb / (ms / reps * 1e-3));
    for (int dbg : {0, 1, 2, 3})
    for (int chunk : {16, 32, 64}) {
        const int grp = 0;
        g_dmma_dbg = dbg;
        printf("dbg=%d (1: no P loads, 2: no P stores) ", dbg);
        launch_cov_update_dmma(P2, ld, n, A, ld, r, sh, panels, n, chunk, 0);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int i = 0; i < reps; i++) launch_cov_update_dmma(P2, ld, n, A, ld, r, sh, panels, n, chunk, 0);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("production kernel, G=%d chunk=%2d   : %8.3f ms  %6.2f TFLOP/s  %7.1f GB/s (incl. panel tiling kernel)\n", grp, chunk, ms / reps, flop / (ms / reps * 1e-3) / 1e12, gb / (ms / reps * 1e-3));
    }
    return 0;
}
