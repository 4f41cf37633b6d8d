// tools/cov_variants.cu — development microbenchmark: template variants of the streaming
// covariance kernel on a synthetic n x n FP64 matrix (values irrelevant to timing).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false \
//        -I conan_slam_b200/csrc tools/cov_variants.cu conan_slam_b200/lib/util.o -o gpurun_out/cov_variants
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cov_update.cuh"

using namespace cslam;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int R, int T, int B, int MINB, int HINT>
float run(double* P, size_t ld, int n, const double* A, size_t lda, int reps) {
    const int nt = (n + T - 1) / T;
    const long long tiles = (long long)nt * (nt + 1) / 2;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_cov_update<R, T, B, MINB, HINT><<<(unsigned)tiles, 256>>>(P, ld, n, A, lda, nt, 0.0, Shard{0, 1});
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; i++) k_cov_update<R, T, B, MINB, HINT><<<(unsigned)tiles, 256>>>(P, ld, n, A, lda, nt, 0.0, Shard{0, 1});
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

template <int M, int T, int B, int MINB, int HINT>
float run_multi(double* P, size_t ld, int n, const double* A, size_t lda, int reps) {
    const int nt = (n + T - 1) / T;
    const long long tiles = (long long)nt * (nt + 1) / 2;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_cov_update_multi<M, T, B, MINB, HINT><<<(unsigned)tiles, 256>>>(P, ld, n, A, lda, nt, Shard{0, 1}, nullptr);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; i++)
        k_cov_update_multi<M, T, B, MINB, HINT><<<(unsigned)tiles, 256>>>(P, ld, n, A, lda, nt, Shard{0, 1}, nullptr);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 40003;
    const int reps = argc > 2 ? atoi(argv[2]) : 10;
    const size_t ld = ((size_t)n + 1 + 15) / 16 * 16;
    double *P, *A;
    CK(cudaMalloc(&P, ld * n * sizeof(double)));
    CK(cudaMalloc(&A, 16 * ld * sizeof(double)));
    CK(cudaMemset(P, 0, ld * n * sizeof(double)));
    CK(cudaMemset(A, 0, 16 * ld * sizeof(double)));
    const double gb = 8.0 * n * ((double)n + 1.0) / 1e9;
#define RUN(R, T, B, MINB, HINT) { float ms = run<R, T, B, MINB, HINT>(P, ld, n, A, ld, reps); \
    printf("n=%d R=%d T=%3d BATCH=%2d MINB=%d HINT=%d : %8.4f ms  %8.1f GB/s\n", n, R, T, B, MINB, HINT, ms, gb / (ms * 1e-3)); }
    RUN(2, 128, 8, 2, 0)
    RUN(2, 128, 8, 2, 1)
    RUN(2, 128, 8, 2, 2)
    RUN(2, 128, 8, 3, 0)
    RUN(2, 128, 8, 4, 0)
    RUN(2, 128, 16, 2, 0)
    RUN(2, 128, 16, 1, 0)
    RUN(2, 128, 4, 4, 0)
    RUN(2, 128, 4, 4, 1)
    RUN(2, 64, 8, 4, 0)
    RUN(2, 64, 8, 4, 1)
    RUN(2, 64, 8, 6, 0)
    RUN(1, 128, 8, 2, 0)
    RUN(1, 128, 8, 4, 1)
#define RUNM(M, T, B, MINB, HINT) { float ms = run_multi<M, T, B, MINB, HINT>(P, ld, n, A, ld, reps); \
    printf("n=%d multi M=%d T=%3d BATCH=%2d MINB=%d HINT=%d : %8.4f ms  %8.1f GB/s  (%.1f updates/ms)\n", n, M, T, B, MINB, HINT, ms, gb / (ms * 1e-3), M / ms); }
    RUNM(2, 128, 4, 4, 1)
    RUNM(2, 128, 8, 4, 1)
    RUNM(3, 128, 4, 4, 1)
    RUNM(4, 128, 4, 3, 1)
    RUNM(4, 128, 4, 4, 1)
    RUNM(4, 128, 4, 5, 1)
    RUNM(4, 128, 8, 3, 1)
    RUNM(4, 128, 8, 4, 1)
    RUNM(4, 128, 2, 5, 1)
    RUNM(4, 64, 4, 4, 1)
    RUNM(4, 64, 8, 4, 1)
    RUNM(8, 128, 4, 4, 1)
    RUNM(8, 128, 8, 4, 1)
    RUNM(8, 128, 4, 3, 1)
    return 0;
}
