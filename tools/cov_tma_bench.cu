// tools/cov_tma_bench.cu — development check + microbenchmark of the TMA / FP64-tensor-core streaming
// covariance pass (csrc/cov_tma.cu) against the FMA pass it replaces (k_cov_update_multi, cov_update.cuh):
// same panels, same P, element-wise comparison over the upper triangle, then CUDA-event timing of both.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -I conan_slam_b200/csrc \
//        tools/cov_tma_bench.cu conan_slam_b200/lib/cov_tma.o conan_slam_b200/lib/util.o -ldl -o tools/bin/cov_tma_bench
//   tools/bin/cov_tma_bench [n=40003] [reps=10] [world=1] [rank=0]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cov_update.cuh"

using namespace cslam;

namespace cslam {
int make_cov_tensor_map(void* out_map64, double* P, size_t ld, size_t rows);
int launch_cov_update_tma(const void* src_map, const void* dst_map, int n, const double* A, size_t lda, int r,
                          double diag_eps, Shard sh, const int* live, int nlive, int num_sms, int stages,
                          cudaStream_t stream, double* dst_ptr = nullptr, size_t dst_ld = 0, int dst_rows = 0);
extern int g_tma_dbg, g_tma_boxr, g_tma_dense, g_tma_direct, g_tma_sub;
}  // namespace cslam

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void k_init(double* P, size_t ld, size_t rows, int n, Shard sh) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * ld) return;
    const size_t lr = idx / ld;
    const int j = (int)(idx % ld);
    const int i = (int)((lr / 128) * sh.world + sh.rank) * 128 + (int)(lr % 128);  // global row of local row lr
    const int d = i > j ? i - j : j - i;
    P[idx] = 2.0 / (1.0 + 0.01 * d) + 1e-3 * sin(0.37 * i + 0.11 * j);
}
__global__ void k_init_panel(double* A, size_t lda, int n, int rows) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)rows * lda) return;
    const int k = (int)(idx / lda), i = (int)(idx % lda);
    A[idx] = i < n ? 0.05 * sin(0.013 * i * (k + 1) + 0.7 * k) + 0.01 * cos(0.001 * i) : 0.0;
}
__device__ unsigned long long* g_where = nullptr;
__global__ void k_compare(const double* P1, const double* P2, size_t ld, size_t rows, int n, Shard sh,
                          unsigned long long* out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * ld) return;
    const size_t lr = idx / ld;
    const int j = (int)(idx % ld);
    const int i = (int)((lr / 128) * sh.world + sh.rank) * 128 + (int)(lr % 128);
    if (i >= n || j >= n || j < i) return;
    const double a = P1[idx], b = P2[idx];
    // relative to the magnitude of the operands (P ~ 1e-2..2, terms ~ 1e-2): results may cancel to ~0
    double rel = fabs(a - b) / fmax(fabs(a), 1e-2);
    if (!(rel == rel)) rel = 1e300;
    atomicMax(out, (unsigned long long)__double_as_longlong(rel));
    if (a != b) {
        const unsigned long long k = atomicAdd(out + 1, 1ULL);
        if (rel > 1e-9 && g_where != nullptr) {  // where the gross mismatches are (first 32)
            const unsigned long long q = atomicAdd(g_where, 1ULL);
            if (q < 32) {
                g_where[1 + 2 * q] = (unsigned long long)i;
                g_where[2 + 2 * q] = (unsigned long long)j;
            }
        }
        (void)k;
    }
}

template <int M>
static void run_ref(double* P, size_t ld, int n, const double* A, size_t lda, Shard sh) {
    const int nt = (n + 127) / 128;
    const long long tiles = shard_tile_count(nt, sh);
    k_cov_update_multi<M, 128, 4, 4, 1><<<(unsigned)tiles, 256>>>(P, ld, n, A, lda, nt, sh, nullptr);
}
static void run_ref_g(int g, double* P, size_t ld, int n, const double* A, size_t lda, Shard sh) {
    switch (g) {
        case 2: run_ref<2>(P, ld, n, A, lda, sh); break;
        case 3: run_ref<3>(P, ld, n, A, lda, sh); break;
        case 4: run_ref<4>(P, ld, n, A, lda, sh); break;
        case 5: run_ref<5>(P, ld, n, A, lda, sh); break;
        case 6: run_ref<6>(P, ld, n, A, lda, sh); break;
        case 7: run_ref<7>(P, ld, n, A, lda, sh); break;
        case 8: run_ref<8>(P, ld, n, A, lda, sh); break;
        default: printf("bad g\n"); exit(1);
    }
}

int main(int argc, char** argv) {
    if (getenv("TMA_BOXR")) g_tma_boxr = atoi(getenv("TMA_BOXR"));
    if (getenv("TMA_DENSE")) g_tma_dense = atoi(getenv("TMA_DENSE"));
    const int n = argc > 1 ? atoi(argv[1]) : 40003;
    const int reps = argc > 2 ? atoi(argv[2]) : 10;
    const Shard sh{argc > 4 ? atoi(argv[4]) : 0, argc > 3 ? atoi(argv[3]) : 1};
    const size_t ld = ((size_t)n + 1 + 15) / 16 * 16;
    const int ntr = (n + 127) / 128;
    size_t owned = 0;
    for (int tr = sh.rank; tr < ntr; tr += sh.world) owned++;
    const size_t rows = sh.world == 1 ? (size_t)n : owned * 128;
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double *P1, *P2, *A;
    unsigned long long* dres;
    CK(cudaMalloc(&P1, ld * rows * sizeof(double)));
    CK(cudaMalloc(&P2, ld * rows * sizeof(double)));
    CK(cudaMalloc(&A, 16 * ld * sizeof(double)));
    CK(cudaMalloc(&dres, 16));
    unsigned char map[128], map1[128];
    if (make_cov_tensor_map(map, P2, ld, rows) != 0) { printf("tensor map: %s\n", cslam_last_error()); return 1; }
    if (make_cov_tensor_map(map1, P1, ld, rows) != 0) { printf("tensor map: %s\n", cslam_last_error()); return 1; }
    const unsigned gi = (unsigned)((rows * ld + 255) / 256);
    k_init_panel<<<(unsigned)((16 * ld + 255) / 256), 256>>>(A, ld, n, 16);
    CK(cudaDeviceSynchronize());
    const double gb = 8.0 * n * ((double)n + 1.0) / 1e9 / sh.world;
    printf("n=%d ld=%zu rows=%zu world=%d rank=%d SMs=%d boxr=%d dense=%d direct=%d sub=%d, algorithmic bytes per pass %.3f GB\n", n, ld, rows, sh.world,
           sh.rank, sms, g_tma_boxr, g_tma_dense, g_tma_direct, g_tma_sub, gb);
    if (getenv("TMA_DET")) {  // run-to-run determinism of each kernel on its own: the same pass twice, bitwise compare
        const int rounds = atoi(getenv("TMA_DET"));
        double* P3 = nullptr;
        CK(cudaMalloc(&P3, ld * rows * sizeof(double)));
        unsigned char map3[128];
        if (make_cov_tensor_map(map3, P3, ld, rows) != 0) return 1;
        unsigned long long* dwhere = nullptr;
        CK(cudaMalloc(&dwhere, 65 * 8));
        CK(cudaMemcpyToSymbol(g_where, &dwhere, sizeof(dwhere)));
        const char* names[] = {"FMA pass", "TMA in place S=5", "TMA in place S=2", "TMA in place S=3", "TMA in place S=4",
                               "TMA out of place S=5", "TMA out of place S=2"};
        for (int it = 0; it < rounds; it++)
            for (int which = 0; which < 7; which++)
                for (int g = 2; g <= 8; g += 2) {
                    k_init<<<gi, 256>>>(P1, ld, rows, n, sh);
                    k_init<<<gi, 256>>>(P2, ld, rows, n, sh);
                    CK(cudaMemset(dres, 0, 16));
                    CK(cudaMemset(dwhere, 0, 65 * 8));
                    const double *ra = P1, *rb = P2;
                    if (which == 0) {
                        run_ref_g(g, P1, ld, n, A, ld, sh);
                        run_ref_g(g, P2, ld, n, A, ld, sh);
                    } else if (which <= 4) {
                        const int stages = which == 1 ? 5 : which == 2 ? 2 : which == 3 ? 3 : 4;
                        launch_cov_update_tma(map1, map1, n, A, ld, 2 * g, 0.0, sh, nullptr, 0, sms, stages, 0, P1, ld, (int)rows);
                        launch_cov_update_tma(map, map, n, A, ld, 2 * g, 0.0, sh, nullptr, 0, sms, stages, 0, P2, ld, (int)rows);
                    } else {
                        const int stages = which == 5 ? 5 : 2;
                        launch_cov_update_tma(map1, map, n, A, ld, 2 * g, 0.0, sh, nullptr, 0, sms, stages, 0, P2, ld, (int)rows);
                        launch_cov_update_tma(map1, map3, n, A, ld, 2 * g, 0.0, sh, nullptr, 0, sms, stages, 0, P3, ld, (int)rows);
                        ra = P2;
                        rb = P3;
                    }
                    CK(cudaDeviceSynchronize());
                    k_compare<<<gi, 256>>>(ra, rb, ld, rows, n, sh, dres);
                    unsigned long long res[2];
                    CK(cudaMemcpy(res, dres, 16, cudaMemcpyDeviceToHost));
                    double rel;
                    memcpy(&rel, &res[0], 8);
                    if (res[1] != 0) {
                        printf("round %d %s g=%d: %llu elements differ between two runs (max rel %.3e)\n", it, names[which], g,
                               res[1], rel);
                        unsigned long long w[65];
                        CK(cudaMemcpy(w, dwhere, sizeof(w), cudaMemcpyDeviceToHost));
                        for (unsigned long long q = 0; q < w[0] && q < 32; q++)
                            printf("   (%llu,%llu) tile (%llu,%llu) in-tile (%llu,%llu)\n", w[1 + 2 * q], w[2 + 2 * q], w[1 + 2 * q] / 128,
                                   w[2 + 2 * q] / 128, w[1 + 2 * q] % 128, w[2 + 2 * q] % 128);
                    }
                }
        printf("determinism check done (%d rounds x 7 variants x 4 ranks)\n", rounds);
        return 0;
    }
    int bad = 0;
    for (int g = 2; g <= 8; g++) {
        k_init<<<gi, 256>>>(P1, ld, rows, n, sh);
        k_init<<<gi, 256>>>(P2, ld, rows, n, sh);
        CK(cudaMemset(dres, 0, 16));
        run_ref_g(g, P1, ld, n, A, ld, sh);
        CK(cudaDeviceSynchronize());
        if (launch_cov_update_tma(map, map, n, A, ld, 2 * g, 0.0, sh, nullptr, 0, sms, 5, 0, P2, ld, (int)rows) != 0) {
            printf("launch: %s\n", cslam_last_error());
            return 1;
        }
        CK(cudaDeviceSynchronize());
        k_compare<<<gi, 256>>>(P1, P2, ld, rows, n, sh, dres);
        unsigned long long res[2];
        CK(cudaMemcpy(res, dres, 16, cudaMemcpyDeviceToHost));
        double rel;
        memcpy(&rel, &res[0], 8);
        printf("g=%d (rank %2d): max rel diff vs FMA pass %.3e, %llu elements differ in the last bits -> %s\n", g, 2 * g,
               rel, res[1], rel < 1e-12 ? "OK" : "MISMATCH");
        if (!(rel < 1e-12)) bad++;
    }
    if (bad) { printf("PARITY FAILED\n"); return 2; }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto time_tma = [&](int g, int stages, int dbg, bool pp = false) {
        g_tma_dbg = dbg;
        float ms = 0.f;
        launch_cov_update_tma(pp ? map1 : map, map, n, A, ld, 2 * g, 0.0, sh, nullptr, 0, sms, stages, 0, P2, ld, (int)rows);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int i = 0; i < reps; i++) launch_cov_update_tma(pp ? map1 : map, map, n, A, ld, 2 * g, 0.0, sh, nullptr, 0, sms, stages, 0, P2, ld, (int)rows);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms, e0, e1));
        g_tma_dbg = 0;
        return ms / reps;
    };
    if (getenv("TMA_ISO")) {  // one pass at a time against back-to-back passes, same direction and alternating (ping-pong)
        const int g = 8, stages = 2;
        auto one = [&](bool fwd) {
            return fwd ? launch_cov_update_tma(map1, map, n, A, ld, 2 * g, 0.0, sh, nullptr, 0, sms, stages, 0, P2, ld, (int)rows)
                       : launch_cov_update_tma(map, map1, n, A, ld, 2 * g, 0.0, sh, nullptr, 0, sms, stages, 0, P1, ld, (int)rows);
        };
        for (int mode = 0; mode < 3; mode++) {
            float tot = 0.f, ms = 0.f;
            one(true);
            CK(cudaDeviceSynchronize());
            if (mode == 0) {  // isolated: synchronise around every launch
                for (int i = 0; i < reps; i++) {
                    CK(cudaEventRecord(e0));
                    one(true);
                    CK(cudaEventRecord(e1));
                    CK(cudaDeviceSynchronize());
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                    tot += ms;
                }
            } else {
                CK(cudaEventRecord(e0));
                for (int i = 0; i < reps; i++) one(mode == 1 ? true : (i & 1) == 0);
                CK(cudaEventRecord(e1));
                CK(cudaDeviceSynchronize());
                CK(cudaEventElapsedTime(&tot, e0, e1));
            }
            printf("n=%d rank 16, 2 stages: %s: %8.4f ms per pass, %7.1f GB/s\n", n,
                   mode == 0 ? "one at a time (sync between)" : mode == 1 ? "back to back, P1 -> P2 every time" : "back to back, ping-pong",
                   tot / reps, gb / (tot / reps * 1e-3));
        }
        return 0;
    }
    for (int g = 2; g <= 8; g++) {
        float ms_ref = 0.f;
        run_ref_g(g, P1, ld, n, A, ld, sh);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int i = 0; i < reps; i++) run_ref_g(g, P1, ld, n, A, ld, sh);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms_ref, e0, e1));
        ms_ref /= reps;
        printf("n=%d g=%d : FMA pass %8.4f ms %7.1f GB/s | TMA+DMMA pass (stages 2..5):", n, g, ms_ref, gb / (ms_ref * 1e-3));
        for (int stages = 2; stages <= 6; stages++) {
            if (stages == 6 && g > 4) continue;
            const float ms = time_tma(g, stages, 0);
            printf("  S%d %7.4f ms %6.1f GB/s", stages, ms, gb / (ms * 1e-3));
        }
        const float ms_pp = time_tma(g, 5, 0, true);
        printf("  | out of place (P1 -> P2) S5 %7.4f ms %6.1f GB/s\n", ms_pp, gb / (ms_pp * 1e-3));
    }
    for (int dbg = 8; dbg <= 24; dbg += 8)
        for (int g = 4; g <= 8; g += 4) {
            const float ms = time_tma(g, 4, dbg);
            const float ms5 = time_tma(g, 5, dbg);
            printf("n=%d g=%d dbg=%d (8 loads without L2 hint, 16 stores without, 24 neither): S4 %8.4f ms %7.1f GB/s  S5 %8.4f ms %7.1f GB/s\n",
                   n, g, dbg, ms, gb / (ms * 1e-3), ms5, gb / (ms5 * 1e-3));
        }
    for (int dbg = 2; dbg <= 6; dbg += 2) {
        for (int stages = 3; stages <= 5; stages += 2) {
            const float ms = time_tma(4, stages, dbg);
            printf("n=%d g=4 stages=%d dbg=%d (2 no stores, 4 no loads, 6 neither): %8.4f ms, %7.1f GB/s of the full pass\n",
                   n, stages, dbg, ms, gb / (ms * 1e-3));
        }
    }
    return 0;
}
