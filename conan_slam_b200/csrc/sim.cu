// sim.cu — on-GPU observation front-end ("next" row 2): the step immediately BEFORE the
// filter hot path in test/main.cpp — Slam::getObservations (slam.h:575-582) = getVisibleLandmarks
// (slam.h:608-683) + computeRangeBearing (slam.h:339-368) — for a simulated world whose landmarks live
// on the device.  On 20k-60k-landmark maps the reference's O(N) host loop (+ a host round trip per scan)
// is what remains serial once the updates themselves run at thousands per second.
//
// Two launches: per-block visible counts, then a stable compaction (landmark order, as the reference's
// loop produces it) with the block offsets recomputed from the counts.  Compiled with -fmad=false and
// written in the oracle's operation order (oracle/slam_oracle.hpp get_observations).
#include <new>

#include "common.cuh"

struct cslam_world {
    int device = 0;
    int n = 0;
    double* lm = nullptr;    // [2][n]
    int* counts = nullptr;   // [blocks]
    double* z = nullptr;     // [2 * n] compacted (range, bearing) pairs
    int* tags = nullptr;     // [n] 1-based landmark tags of the visible ones, ascending
    int* m = nullptr;        // number visible
    // known-association bookkeeping on the device (EKF.cpp:146-233, the reference's mTABLE)
    int* table = nullptr;    // [n] map slot (1-based) of every world landmark, 0 = not in the map yet
    double* zf = nullptr;    // [2 * n] observations of landmarks already in the map
    double* zn = nullptr;    // [2 * n] observations of new landmarks
    int* idf = nullptr;      // [n] map slots of zf
    int* cnt = nullptr;      // [2] mf, mn
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    cudaStream_t stream = nullptr;
};

namespace cslam {

constexpr int kVisThreads = 256;

__device__ __forceinline__ bool visible(double lx, double ly, double x, double y, double phi, double rmax,
                                        double& dx, double& dy) {
    dx = lx - x;
    dy = ly - y;
    // slam.h:644-648: bounding box, forward half-plane, range circle
    return (fabs(dx) < rmax && fabs(dy) < rmax) && ((dx * cos(phi) + dy * sin(phi)) > 0.0) &&
           ((dx * dx + dy * dy) < rmax * rmax);
}

__global__ void __launch_bounds__(kVisThreads) k_vis_count(const double* __restrict__ lm, int n, double x, double y,
                                                           double phi, double rmax, int* __restrict__ counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double dx, dy;
    const bool v = i < n && visible(lm[i], lm[(size_t)n + i], x, y, phi, rmax, dx, dy);
    const int c = __syncthreads_count(v ? 1 : 0);
    if (threadIdx.x == 0) counts[blockIdx.x] = c;
}

__global__ void __launch_bounds__(kVisThreads) k_vis_compact(const double* __restrict__ lm, int n, double x, double y,
                                                             double phi, double rmax,
                                                             const int* __restrict__ counts, double* __restrict__ z,
                                                             int* __restrict__ tags, int* __restrict__ m) {
    __shared__ int s_part[kVisThreads];
    __shared__ int s_warp[kVisThreads / 32];
    // exclusive offset of this block = sum of the counts of the blocks before it
    int acc = 0;
    for (int b = threadIdx.x; b < (int)blockIdx.x; b += blockDim.x) acc += counts[b];
    s_part[threadIdx.x] = acc;
    __syncthreads();
    for (int s = kVisThreads / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) s_part[threadIdx.x] += s_part[threadIdx.x + s];
        __syncthreads();
    }
    const int base = s_part[0];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double dx = 0.0, dy = 0.0;
    const bool v = i < n && visible(lm[i], lm[(size_t)n + i], x, y, phi, rmax, dx, dy);
    const unsigned ballot = __ballot_sync(0xffffffffu, v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_warp[warp] = __popc(ballot);
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; w++) woff += s_warp[w];
    if (v) {
        const int k = base + woff + __popc(ballot & ((1u << lane) - 1u));
        z[2 * (size_t)k] = sqrt(dx * dx + dy * dy);   // slam.h:352
        z[2 * (size_t)k + 1] = atan2(dy, dx) - phi;   // slam.h:353 (bearing not wrapped)
        tags[k] = i + 1;
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
        int tot = base;
        for (int w = 0; w < kVisThreads / 32; w++) tot += s_warp[w];
        *m = tot;
    }
}

// EKF.cpp:146-233 dataAssociateTable for the visible list left on the device by k_vis_compact: landmark
// `tag` is known when table[tag - 1] != 0 (then idf = that slot), else new — new ones receive the slots
// nf + 1, nf + 2, ... in list order (EKF.cpp:212-226) and the table is updated.  One CTA, stable
// compaction in chunks of blockDim.x with running offsets.
__global__ void __launch_bounds__(1024) k_table_associate(const double* __restrict__ z, const int* __restrict__ tags,
                                                          const int* __restrict__ m_dev, int* __restrict__ table,
                                                          int nf, double* __restrict__ zf, int* __restrict__ idf,
                                                          double* __restrict__ zn, int* __restrict__ cnt) {
    __shared__ int s_known[32], s_new[32];
    __shared__ int base_known, base_new;
    const int m = *m_dev;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) { base_known = 0; base_new = 0; }
    __syncthreads();
    for (int c0 = 0; c0 < m; c0 += blockDim.x) {
        const int k = c0 + threadIdx.x;
        const bool in = k < m;
        const int tag = in ? tags[k] : 0;
        const int slot = in ? table[tag - 1] : 0;
        const bool known = in && slot != 0, fresh = in && slot == 0;
        const unsigned bk = __ballot_sync(0xffffffffu, known), bn = __ballot_sync(0xffffffffu, fresh);
        if (lane == 0) { s_known[warp] = __popc(bk); s_new[warp] = __popc(bn); }
        __syncthreads();
        int ok = base_known, on = base_new;
        for (int w = 0; w < warp; w++) { ok += s_known[w]; on += s_new[w]; }
        const unsigned below = (1u << lane) - 1u;
        if (known) {
            const int p = ok + __popc(bk & below);
            zf[2 * (size_t)p] = z[2 * (size_t)k];
            zf[2 * (size_t)p + 1] = z[2 * (size_t)k + 1];
            idf[p] = slot;
        }
        if (fresh) {
            const int p = on + __popc(bn & below);
            zn[2 * (size_t)p] = z[2 * (size_t)k];
            zn[2 * (size_t)p + 1] = z[2 * (size_t)k + 1];
            table[tag - 1] = nf + p + 1;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 0; w < nw; w++) { base_known += s_known[w]; base_new += s_new[w]; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { cnt[0] = base_known; cnt[1] = base_new; }
}

}  // namespace cslam

using namespace cslam;

extern "C" {

int cslam_world_create(cslam_world_t** out, const double* landmarks_2xN, int num_landmarks, int device) {
    CSLAM_NVTX_RANGE();
    CSLAM_REQUIRE(out != nullptr && num_landmarks >= 0 && (landmarks_2xN != nullptr || num_landmarks == 0),
                  CSLAM_ERR_BAD_ARG, "bad argument");
    *out = nullptr;
    int count = 0;
    CSLAM_CUDA(cudaGetDeviceCount(&count));
    CSLAM_REQUIRE(device >= 0 && device < count, CSLAM_ERR_CUDA, "no such CUDA device (no CPU fallback exists)");
    CSLAM_CUDA(cudaSetDevice(device));
    cslam_world* w = new (std::nothrow) cslam_world();
    CSLAM_REQUIRE(w != nullptr, CSLAM_ERR_BAD_ARG, "out of host memory");
    w->device = device;
    w->n = num_landmarks;
    const int n1 = num_landmarks > 0 ? num_landmarks : 1;
    const int blocks = (n1 + kVisThreads - 1) / kVisThreads;
    w->pinned_bytes = 64 + (size_t)n1 * (4 * sizeof(double) + 2 * sizeof(int));
    bool ok = cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaMalloc(&w->lm, 2 * (size_t)n1 * sizeof(double)) == cudaSuccess &&
              cudaMalloc(&w->counts, blocks * sizeof(int)) == cudaSuccess &&
              cudaMalloc(&w->z, 2 * (size_t)n1 * sizeof(double)) == cudaSuccess &&
              cudaMalloc(&w->tags, (size_t)n1 * sizeof(int)) == cudaSuccess &&
              cudaMalloc(&w->m, sizeof(int)) == cudaSuccess &&
              cudaMalloc(&w->table, (size_t)n1 * sizeof(int)) == cudaSuccess &&
              cudaMemset(w->table, 0, (size_t)n1 * sizeof(int)) == cudaSuccess &&
              cudaMalloc(&w->zf, 2 * (size_t)n1 * sizeof(double)) == cudaSuccess &&
              cudaMalloc(&w->zn, 2 * (size_t)n1 * sizeof(double)) == cudaSuccess &&
              cudaMalloc(&w->idf, (size_t)n1 * sizeof(int)) == cudaSuccess &&
              cudaMalloc(&w->cnt, 2 * sizeof(int)) == cudaSuccess &&
              cudaMallocHost(&w->pinned, w->pinned_bytes) == cudaSuccess;
    if (ok && num_landmarks > 0) {
        // the reference's LM is 2 x N column-major (x_i, y_i interleaved); the device keeps [2][N]
        double* stage = static_cast<double*>(malloc(2 * (size_t)num_landmarks * sizeof(double)));
        ok = stage != nullptr;
        if (ok) {
            for (int i = 0; i < num_landmarks; i++) {
                stage[i] = landmarks_2xN[2 * (size_t)i];
                stage[(size_t)num_landmarks + i] = landmarks_2xN[2 * (size_t)i + 1];
            }
            ok = cudaMemcpy(w->lm, stage, 2 * (size_t)num_landmarks * sizeof(double), cudaMemcpyHostToDevice) ==
                 cudaSuccess;
            free(stage);
        }
    }
    if (!ok) {
        set_last_error("cslam_world_create: allocation or upload failed (%s)", cudaGetErrorString(cudaGetLastError()));
        cslam_world_destroy(w);
        return CSLAM_ERR_CUDA;
    }
    *out = w;
    return CSLAM_OK;
}

int cslam_world_destroy(cslam_world_t* w) {
    CSLAM_NVTX_RANGE();
    if (!w) return CSLAM_OK;
    cudaSetDevice(w->device);
    if (w->stream) cudaStreamSynchronize(w->stream);
    cudaFree(w->lm); cudaFree(w->counts); cudaFree(w->z); cudaFree(w->tags); cudaFree(w->m);
    cudaFree(w->table); cudaFree(w->zf); cudaFree(w->zn); cudaFree(w->idf); cudaFree(w->cnt);
    if (w->pinned) cudaFreeHost(w->pinned);
    if (w->stream) cudaStreamDestroy(w->stream);
    delete w;
    return CSLAM_OK;
}

int cslam_world_observe(cslam_world_t* w, const double x_true[3], double max_range, int max_out, double* Z,
                        int32_t* tags, int* m_out) {
    CSLAM_NVTX_RANGE();
    CSLAM_REQUIRE(w != nullptr && x_true != nullptr && m_out != nullptr && max_out >= 0, CSLAM_ERR_BAD_ARG,
                  "bad argument");
    CSLAM_REQUIRE(max_out == 0 || (Z != nullptr && tags != nullptr), CSLAM_ERR_BAD_ARG, "null output");
    CSLAM_CUDA(cudaSetDevice(w->device));
    *m_out = 0;
    if (w->n == 0) return CSLAM_OK;
    const int blocks = (w->n + kVisThreads - 1) / kVisThreads;
    count_launch();
    k_vis_count<<<blocks, kVisThreads, 0, w->stream>>>(w->lm, w->n, x_true[0], x_true[1], x_true[2], max_range,
                                                       w->counts);
    count_launch();
    k_vis_compact<<<blocks, kVisThreads, 0, w->stream>>>(w->lm, w->n, x_true[0], x_true[1], x_true[2], max_range,
                                                         w->counts, w->z, w->tags, w->m);
    CSLAM_CUDA(cudaGetLastError());
    const int cap = max_out < w->n ? max_out : w->n;
    char* pin = static_cast<char*>(w->pinned);
    int* pm = reinterpret_cast<int*>(pin);
    double* pz = reinterpret_cast<double*>(pin + 16);
    int* pt = reinterpret_cast<int*>(pin + 16 + 2 * (size_t)cap * sizeof(double));
    CSLAM_CUDA(cudaMemcpyAsync(pm, w->m, sizeof(int), cudaMemcpyDeviceToHost, w->stream));
    if (cap > 0) {
        CSLAM_CUDA(cudaMemcpyAsync(pz, w->z, 2 * (size_t)cap * sizeof(double), cudaMemcpyDeviceToHost, w->stream));
        CSLAM_CUDA(cudaMemcpyAsync(pt, w->tags, (size_t)cap * sizeof(int), cudaMemcpyDeviceToHost, w->stream));
    }
    CSLAM_CUDA(cudaStreamSynchronize(w->stream));
    const int m = *pm;
    *m_out = m;  // the true count, also when it exceeds max_out (the caller then sees a truncated list)
    const int take = m < cap ? m : cap;
    for (int k = 0; k < take; k++) {
        Z[2 * k] = pz[2 * k];
        Z[2 * k + 1] = pz[2 * k + 1];
        tags[k] = pt[k];
    }
    return CSLAM_OK;
}

// Slam::getObservations (slam.h:575-582) followed by Slam::dataAssociateTable (slam.h:454-457 -> EKF.cpp:146-233)
// — test/main.cpp:177-186 — without leaving the device in between: visibility, range / bearing, the lookup in
// the association table (kept on the device, the reference's mTABLE) and the split into known (ZF, idf) and
// new (ZN) observations; ONE read-back brings the two compact lists to the host, which then calls
// cslam_ekf_update / _augment (or cslam_ekf_observe_step).  num_map_landmarks = (X.rows() - 3) / 2, the `nf`
// of EKF.cpp:212.  Outputs hold at most max_out observations each; *mf / *mn are the true counts.
int cslam_world_observe_associate(cslam_world_t* w, const double x_true[3], double max_range, int num_map_landmarks,
                                  int max_out, double* ZF, int32_t* idf, int* mf, double* ZN, int* mn) {
    CSLAM_NVTX_RANGE();
    CSLAM_REQUIRE(w != nullptr && x_true != nullptr && mf != nullptr && mn != nullptr && max_out >= 0 &&
                      num_map_landmarks >= 0,
                  CSLAM_ERR_BAD_ARG, "bad argument");
    CSLAM_REQUIRE(max_out == 0 || (ZF && idf && ZN), CSLAM_ERR_BAD_ARG, "null output");
    CSLAM_CUDA(cudaSetDevice(w->device));
    *mf = *mn = 0;
    if (w->n == 0) return CSLAM_OK;
    const int blocks = (w->n + kVisThreads - 1) / kVisThreads;
    count_launch();
    k_vis_count<<<blocks, kVisThreads, 0, w->stream>>>(w->lm, w->n, x_true[0], x_true[1], x_true[2], max_range,
                                                       w->counts);
    count_launch();
    k_vis_compact<<<blocks, kVisThreads, 0, w->stream>>>(w->lm, w->n, x_true[0], x_true[1], x_true[2], max_range,
                                                         w->counts, w->z, w->tags, w->m);
    count_launch();
    k_table_associate<<<1, 1024, 0, w->stream>>>(w->z, w->tags, w->m, w->table, num_map_landmarks, w->zf, w->idf, w->zn,
                                                 w->cnt);
    CSLAM_CUDA(cudaGetLastError());
    const int cap = max_out < w->n ? max_out : w->n;
    char* pin = static_cast<char*>(w->pinned);
    int* pc = reinterpret_cast<int*>(pin);
    double* pzf = reinterpret_cast<double*>(pin + 64);
    double* pzn = pzf + 2 * (size_t)cap;
    int* pid = reinterpret_cast<int*>(pzn + 2 * (size_t)cap);
    CSLAM_CUDA(cudaMemcpyAsync(pc, w->cnt, 2 * sizeof(int), cudaMemcpyDeviceToHost, w->stream));
    if (cap > 0) {
        CSLAM_CUDA(cudaMemcpyAsync(pzf, w->zf, 2 * (size_t)cap * sizeof(double), cudaMemcpyDeviceToHost, w->stream));
        CSLAM_CUDA(cudaMemcpyAsync(pzn, w->zn, 2 * (size_t)cap * sizeof(double), cudaMemcpyDeviceToHost, w->stream));
        CSLAM_CUDA(cudaMemcpyAsync(pid, w->idf, (size_t)cap * sizeof(int), cudaMemcpyDeviceToHost, w->stream));
    }
    CSLAM_CUDA(cudaStreamSynchronize(w->stream));
    *mf = pc[0];
    *mn = pc[1];
    const int tf = pc[0] < cap ? pc[0] : cap, tn = pc[1] < cap ? pc[1] : cap;
    for (int k = 0; k < tf; k++) { ZF[2 * k] = pzf[2 * k]; ZF[2 * k + 1] = pzf[2 * k + 1]; idf[k] = pid[k]; }
    for (int k = 0; k < tn; k++) { ZN[2 * k] = pzn[2 * k]; ZN[2 * k + 1] = pzn[2 * k + 1]; }
    return CSLAM_OK;
}

// The association table (mTABLE, slam.h:105): read it back / clear it (a new run on the same world).
int cslam_world_get_table(cslam_world_t* w, int32_t* table) {
    CSLAM_NVTX_RANGE();
    CSLAM_REQUIRE(w != nullptr && (table != nullptr || w->n == 0), CSLAM_ERR_BAD_ARG, "bad argument");
    CSLAM_CUDA(cudaSetDevice(w->device));
    if (w->n == 0) return CSLAM_OK;
    CSLAM_CUDA(cudaMemcpyAsync(table, w->table, (size_t)w->n * sizeof(int), cudaMemcpyDeviceToHost, w->stream));
    CSLAM_CUDA(cudaStreamSynchronize(w->stream));
    return CSLAM_OK;
}
int cslam_world_reset_table(cslam_world_t* w) {
    CSLAM_NVTX_RANGE();
    CSLAM_REQUIRE(w != nullptr, CSLAM_ERR_BAD_ARG, "bad argument");
    CSLAM_CUDA(cudaSetDevice(w->device));
    if (w->n > 0) CSLAM_CUDA(cudaMemsetAsync(w->table, 0, (size_t)w->n * sizeof(int), w->stream));
    return CSLAM_OK;
}

}  // extern "C"
