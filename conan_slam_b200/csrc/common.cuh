// common.cuh — internals shared by the sm_100a kernels and the C-ABI glue of libcslam.so.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX 3: ranges cost nothing unless a tool (nsys / ncu) is attached

#include "../../include/cslam.h"
#include "shard_map.h"

namespace cslam {

void set_last_error(const char* fmt, ...);
void count_launch();  // every kernel launch of the library is counted (bench.py: gpu_launches)

// One NVTX range per C-ABI call (the reference has no tracing at all): timelines show
// cslam_ekf_scan / cslam_pf_resample / ... around the kernels they launch.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};
#define CSLAM_NVTX_RANGE() ::cslam::NvtxRange nvtx_range__(__func__)

#define CSLAM_CUDA(call)                                                                        \
    do {                                                                                        \
        cudaError_t err__ = (call);                                                             \
        if (err__ != cudaSuccess) {                                                             \
            ::cslam::set_last_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,                \
                                    cudaGetErrorString(err__));                                 \
            return CSLAM_ERR_CUDA;                                                              \
        }                                                                                       \
    } while (0)

#define CSLAM_REQUIRE(cond, code, msg)                                                          \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            ::cslam::set_last_error("%s: %s", __func__, msg);                                   \
            return code;                                                                        \
        }                                                                                       \
    } while (0)

constexpr double kPi = 3.14159265358979323846;
constexpr double kFltMin = 1.17549435082228750797e-38;  // std::numeric_limits<float>::min(), slam.h:719

// slam.h:816-829 pi2Pi (FP64 restatement: fmod by 2*pi, then one correction each side)
__host__ __device__ __forceinline__ double pi2pi(double a) {
    a = fmod(a, 2.0 * kPi);
    if (a > kPi) a = a - 2.0 * kPi;
    if (a < -kPi) a = a + 2.0 * kPi;
    return a;
}

// Symmetric read of the upper-triangle-authoritative covariance: P(i,j) == P(j,i).
__device__ __forceinline__ double psym(const double* __restrict__ P, size_t ld, int i, int j) {
    return i <= j ? P[(size_t)i * ld + j] : P[(size_t)j * ld + i];
}

// Linearised range-bearing observation of one landmark: EKF.cpp:354-404 / PF.cpp:97-127.
// hu = d h / d(x,y,phi) (2x3), lu = d h / d(lx,ly) (2x2); zb is NOT wrapped.
struct ObsLin {
    double zr, zb;
    double hu[2][3];
    double lu[2][2];
};
__device__ __forceinline__ ObsLin observe_lin(double x, double y, double phi, double lx, double ly) {
    ObsLin o;
    const double dx = lx - x, dy = ly - y;
    const double d2 = dx * dx + dy * dy;
    const double d = sqrt(d2);
    const double xd = dx / d, yd = dy / d, xd2 = dx / d2, yd2 = dy / d2;
    o.zr = d;
    o.zb = atan2(dy, dx) - phi;
    o.hu[0][0] = -xd;  o.hu[0][1] = -yd;  o.hu[0][2] = 0.0;
    o.hu[1][0] = yd2;  o.hu[1][1] = -xd2; o.hu[1][2] = -1.0;
    o.lu[0][0] = xd;   o.lu[0][1] = yd;
    o.lu[1][0] = -yd2; o.lu[1][1] = xd2;
    return o;
}

// 128-bit global accessors for the streaming covariance kernels.
__device__ __forceinline__ double2 ld128(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st128(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }

// ------------------------------------------------------------------------------------
// Small dense algebra in the oracle's operation order
// ------------------------------------------------------------------------------------
template <int R, int K, int Cn>
__device__ __forceinline__ void mm(const double (&A)[R][K], const double (&B)[K][Cn], double (&C)[R][Cn]) {
#pragma unroll
    for (int i = 0; i < R; i++)
#pragma unroll
        for (int j = 0; j < Cn; j++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < K; k++) s += A[i][k] * B[k][j];
            C[i][j] = s;
        }
}
template <int R, int Cn>
__device__ __forceinline__ void tr(const double (&A)[R][Cn], double (&B)[Cn][R]) {
#pragma unroll
    for (int i = 0; i < R; i++)
#pragma unroll
        for (int j = 0; j < Cn; j++) B[j][i] = A[i][j];
}

// oracle::PartialPivLU<T>::inverse(), unrolled for N in {1,2,3} with register-only row swaps
template <int N>
__device__ __forceinline__ void inv_lu(const double (&M)[N][N], double (&inv)[N][N]) {
    double lu[N][N];
    int perm[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        perm[i] = i;
#pragma unroll
        for (int j = 0; j < N; j++) lu[i][j] = M[i][j];
    }
#pragma unroll
    for (int k = 0; k < N; k++) {
        int piv = k;
        double best = fabs(lu[k][k]);
#pragma unroll
        for (int i = k + 1; i < N; i++)
            if (fabs(lu[i][k]) > best) { best = fabs(lu[i][k]); piv = i; }
#pragma unroll
        for (int i = k + 1; i < N; i++)
            if (piv == i) {
#pragma unroll
                for (int j = 0; j < N; j++) { const double t = lu[k][j]; lu[k][j] = lu[i][j]; lu[i][j] = t; }
                const int tp = perm[k]; perm[k] = perm[i]; perm[i] = tp;
            }
#pragma unroll
        for (int i = k + 1; i < N; i++) {
            lu[i][k] = lu[i][k] / lu[k][k];
            const double f = lu[i][k];
#pragma unroll
            for (int j = k + 1; j < N; j++) lu[i][j] -= f * lu[k][j];
        }
    }
#pragma unroll
    for (int col = 0; col < N; col++) {
        double y[N];
#pragma unroll
        for (int i = 0; i < N; i++) {
            double s = (perm[i] == col) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < i; k++) s -= lu[i][k] * y[k];
            y[i] = s;
        }
#pragma unroll
        for (int i = N - 1; i >= 0; i--) {
            double s = y[i];
#pragma unroll
            for (int k = i + 1; k < N; k++) s -= lu[i][k] * inv[k][col];
            inv[i][col] = s / lu[i][i];
        }
    }
}

// oracle::cholesky_decomposition (slam.h:413-436): LLT lower; a non-positive pivot would send
// the reference to its eigen-solver branch (non-unique factor) — here it yields
// the zero matrix, which every caller turns into a skipped/degenerate result, and is counted.
template <int N>
__device__ __forceinline__ bool chol_lower(const double (&M)[N][N], double (&L)[N][N]) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j < N; j++) L[i][j] = 0.0;
#pragma unroll
    for (int j = 0; j < N; j++) {
        double d = M[j][j];
#pragma unroll
        for (int k = 0; k < j; k++) d -= L[j][k] * L[j][k];
        if (!(d > 0.0)) ok = false;
        const double ljj = sqrt(d);
        L[j][j] = ljj;
#pragma unroll
        for (int i = j + 1; i < N; i++) {
            double s = M[i][j];
#pragma unroll
            for (int k = 0; k < j; k++) s -= L[i][k] * L[j][k];
            L[i][j] = s / ljj;
        }
    }
    bool fin = true;
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j < N; j++) fin = fin && isfinite(L[i][j]);
    if (!ok || !fin) {
#pragma unroll
        for (int i = 0; i < N; i++)
#pragma unroll
            for (int j = 0; j < N; j++) L[i][j] = 0.0;
        return false;
    }
    return true;
}

// Pinned staging + device scratch shared by the handle types.
struct Staging {
    void* host = nullptr;  // pinned
    void* dev = nullptr;
    size_t bytes = 0;
};

}  // namespace cslam
