// ekf_dmma.cu — joint (batch) covariance update  P <- P - W1 W1^T  (slam.h:260) for a
// rank-r panel (r = 2m <= 64) on the FP64 tensor cores of sm_100a.
//
// FP64 is not a tcgen05.mma kind: Blackwell's FP64 tensor path is the warp-level
// mma.sync.aligned.m8n8k4.f64 (SASS: DMMA.8x8x4; the wider PTX shapes m16n8k{4,8,16} lower to
// the same instruction).  The update is a real rank-r contraction (8 flop/B at r = 64, above
// the FP64 ridge) only in this batched case; the sequential rank-1/2 updates stay on the
// streaming kernel in cov_update.cuh.
//
// Structure (one CTA per SM, 8 warps, each a 32 x 16 sub-tile = 4 x 2 DMMA blocks):
//   * a CTA owns a CHUNK of consecutive 128 x 32 tiles of ONE 128-row strip of the upper
//     triangle; the strip's (negated) row panel is brought into shared memory once per chunk;
//   * the column panel of every tile arrives by a single TMA bulk copy (cp.async.bulk +
//     mbarrier complete_tx) from a pre-tiled, pre-padded copy of the panel into a 4-deep ring;
//     full/empty mbarriers hand the buffers between the producer lane and the 8 warps, so
//     there is no CTA-wide barrier in the loop and the warps drift apart (one warp's load/store
//     phase overlaps another's DMMA phase);
//   * the accumulator fragments ARE the covariance: every thread stages the fragments of tile
//     t+2 global -> shared with 16-byte cp.async copies while tiles t and t+1 compute, reads them
//     into registers just in time (C operand of the first k-step) and stores the result straight
//     from the accumulators; P is read once and written once;
//   * panels sit in shared memory with a row stride = 4 (mod 16) doubles: conflict-free
//     fragment reads (the 16 lanes of a half-warp cover 16 distinct 8-byte banks).
#include "common.cuh"
#include "cov_update.cuh"

namespace cslam {

constexpr int DM_TM = 128, DM_TN = 32, DM_K = 64;
constexpr int DM_SR = DM_TM + 4;  // 132: row-panel stride  (doubles)
constexpr int DM_SC = DM_TN + 4;  // 36 : column-panel stride
constexpr int DM_NBUF = 4;        // column-panel ring depth
constexpr int DM_THREADS = 8 * 32;  // 8 warps, one 32 x 16 sub-tile each
constexpr int DM_STAGE = 8 * DM_THREADS * 16;  // bytes of one covariance staging buffer (8 x 16 B per thread)
constexpr int DM_SMEM = (DM_K * DM_SR + DM_NBUF * DM_K * DM_SC) * (int)sizeof(double) + 2 * DM_STAGE + 128;

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
// Ampere-style async copy global -> shared, 16 bytes, L2 only; src_bytes = 0 zero-fills without
// touching global memory (masked elements of diagonal / edge tiles).
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// Pre-tiled panels: Ac[tc][k][DM_SC] = A[k][tc*DM_TN + jj], Ar[tr][k][DM_SR] = -A[k][tr*DM_TM + ii]
// (zero beyond n / beyond rank r), each tile a contiguous block one bulk copy brings in.
__global__ void __launch_bounds__(256) k_dmma_panels(const double* __restrict__ A, size_t lda, int n, int r,
                                                     int ncols, double* __restrict__ Ar, double* __restrict__ Ac) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (i >= ncols) return;
    const double v = (k < r && i < n) ? A[(size_t)k * lda + i] : 0.0;
    Ac[(size_t)(i / DM_TN) * (DM_K * DM_SC) + k * DM_SC + (i % DM_TN)] = v;
    Ar[(size_t)(i / DM_TM) * (DM_K * DM_SR) + k * DM_SR + (i % DM_TM)] = -v;
}

// One 32 x 16 warp sub-tile (4 x 2 DMMA blocks): f += (-row panel) x (column panel) over the
// k-steps [ks0, ks1) (4 k each); the a/b fragments of k-step s+1 are fetched from shared memory
// while the 8 DMMAs of k-step s issue.
__device__ __forceinline__ void dmma_ksteps(double (&f)[4][2][2], const double* __restrict__ pr,
                                            const double* __restrict__ pc, int ks0, int ks1) {
    double a0[4], b0[2], a1[4], b1[2];
#pragma unroll
    for (int x = 0; x < 4; x++) a0[x] = pr[ks0 * 4 * DM_SR + x * 8];
#pragma unroll
    for (int x = 0; x < 2; x++) b0[x] = pc[ks0 * 4 * DM_SC + x * 8];
#pragma unroll
    for (int ks = ks0; ks < DM_K / 4; ks += 2) {
        if (ks >= ks1) break;
        const bool odd = ks + 1 < ks1;
        if (odd) {
#pragma unroll
            for (int x = 0; x < 4; x++) a1[x] = pr[(ks + 1) * 4 * DM_SR + x * 8];
#pragma unroll
            for (int x = 0; x < 2; x++) b1[x] = pc[(ks + 1) * 4 * DM_SC + x * 8];
        }
#pragma unroll
        for (int bi = 0; bi < 4; bi++)
#pragma unroll
            for (int bj = 0; bj < 2; bj++) dmma884(f[bi][bj][0], f[bi][bj][1], a0[bi], b0[bj]);
        if (odd) {
            if (ks + 2 < ks1) {
#pragma unroll
                for (int x = 0; x < 4; x++) a0[x] = pr[(ks + 2) * 4 * DM_SR + x * 8];
#pragma unroll
                for (int x = 0; x < 2; x++) b0[x] = pc[(ks + 2) * 4 * DM_SC + x * 8];
            }
#pragma unroll
            for (int bi = 0; bi < 4; bi++)
#pragma unroll
                for (int bj = 0; bj < 2; bj++) dmma884(f[bi][bj][0], f[bi][bj][1], a1[bi], b1[bj]);
        }
    }
}

template <bool FULL>
__global__ void __launch_bounds__(DM_THREADS, 1) k_cov_update_dmma(double* P, size_t ld, int n,
                                                                   const double* __restrict__ Ar,
                                                                   const double* __restrict__ Ac, int rp, int nbc,
                                                                   int chunk, Shard sh, int dbg, ptrdiff_t dshift,
                                                                   double diag_eps) {
    // dshift: distance (in doubles) from the array that is READ (P) to the array that is WRITTEN — 0 in place, the
    // ping-pong twin for the deferred passes of ekf_lazy.cuh.  diag_eps: added to every diagonal element (the
    // heading terms of a bank, slam.h:719).
#ifndef CSLAM_DMMA_ABLATION
    dbg = 0;  // the load/store ablation of tools/dmma_bench.cu is compiled out of the library
#endif
    extern __shared__ __align__(128) double smem[];
    double* sR = smem;                 // [rp][DM_SR]  negated row panel of the strip
    double* sC = smem + DM_K * DM_SR;  // [DM_NBUF][rp][DM_SC] column panels (ring)
    char* sP = reinterpret_cast<char*>(sC + DM_NBUF * DM_K * DM_SC);  // [2][8][DM_THREADS] x 16 B staging
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * DM_STAGE);

    const int tr = blockIdx.y * sh.world + sh.rank;                 // 128-row strip (global tile row)
    const int c0 = (DM_TM / DM_TN) * tr + blockIdx.x * chunk;       // first 32-column tile of this chunk
    if (c0 >= nbc) return;
    const int c1 = min(c0 + chunk, nbc);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * DM_NBUF, bar_row = bar_empty + 8 * DM_NBUF;
    const uint32_t row_bytes = (uint32_t)rp * DM_SR * sizeof(double);
    const uint32_t col_bytes = (uint32_t)rp * DM_SC * sizeof(double);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < DM_NBUF; s++) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 8);  // one arrival per warp
        }
        mbar_init(bar_row, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // Producer role (lane 0 of warp 0, inside its own consumer loop): at step q it refills the ring
    // slot that tile q-2 used with the panel of tile q+2, so it only ever waits for warps that lag
    // two whole tiles behind.
    auto produce = [&](int q) {  // q = tile index within the chunk
        if (c0 + q >= c1) return;
        const int s = q % DM_NBUF;
        if (q >= DM_NBUF) mbar_wait(bar_empty + 8 * s, ((q / DM_NBUF) - 1) & 1);
        mbar_expect_tx(bar_full + 8 * s, col_bytes);
        bulk_g2s(smem_u32(sC + s * (DM_K * DM_SC)), Ac + (size_t)(c0 + q) * (DM_K * DM_SC), col_bytes,
                 bar_full + 8 * s);
    };
    if (tid == 0) {
        mbar_expect_tx(bar_row, row_bytes);
        bulk_g2s(smem_u32(sR), Ar + (size_t)tr * (DM_K * DM_SR), row_bytes, bar_row);
        produce(0);
        produce(1);
    }

    const int wr = warp >> 1, wc = warp & 1;
    const int lr = lane >> 2, lc = lane & 3;
    const int iw = tr * DM_TM + wr * 32;
    // row bases of the 4 fragment rows this lane owns (rows >= n are never dereferenced)
    double* prow[4];
#pragma unroll
    for (int bi = 0; bi < 4; bi++) prow[bi] = P + shard_lrow(sh, iw + bi * 8 + lr) * ld + 2 * lc;
    const uint32_t stage0 = smem_u32(sP) + tid * 16;  // this thread's slots: stage0 + buf*DM_STAGE + frag*4096

    // sub-tile state: 0 = nothing to do (below the diagonal / outside), 1 = interior (no masks), 2 = masked
    auto tile_kind = [&](int tc) -> int {
        if (tc >= c1) return 0;
        const int jw = tc * DM_TN + wc * 16;
        if (!((jw + 15 >= iw) && (iw < n) && (jw < n))) return 0;
        return (jw >= iw + 32 && jw + 16 <= n && iw + 32 <= n) ? 1 : 2;
    };
    // Covariance fragments travel global -> shared by per-thread async copies (each thread stages
    // exactly the 8 x 16 B it will consume, so there is no cross-thread hazard and no bank
    // conflict), NOT by register loads: ptxas tracks DMMA results, LDS and LDG on the same six
    // scoreboard counters per warp, and a register load shares a counter with the panel LDS — every
    // DMMA that waits for its a/b fragments would then also wait for the newest global loads.
    auto fetch_tile = [&](int tc, int kind, int buf) {
        if (kind && !(dbg & 1)) {
            const int jw = tc * DM_TN + wc * 16;
            const uint32_t dst = stage0 + buf * DM_STAGE;
#pragma unroll
            for (int bi = 0; bi < 4; bi++) {
                const int i = iw + bi * 8 + lr;
#pragma unroll
                for (int bj = 0; bj < 2; bj++) {
                    const int j = jw + bj * 8 + 2 * lc;
                    const bool ok = kind == 1 || (i < n && j < n && j + 1 >= i);
                    cp_async16(dst + (bi * 2 + bj) * (DM_THREADS * 16), ok ? prow[bi] + jw + bj * 8 : P, ok ? 16u : 0u);
                }
            }
        }
        cp_async_commit();  // always: keeps the group count per tile fixed
    };

    double f[4][2][2];
    const double* pr = sR + lc * DM_SR + wr * 32 + lr;
    const int nks = FULL ? DM_K / 4 : rp / 4;
    fetch_tile(c0, tile_kind(c0), 0);
    fetch_tile(c0 + 1, tile_kind(c0 + 1), 1);
    mbar_wait(bar_row, 0);
    for (int t = c0; t < c1; t++) {
        const int q = t - c0, s = q % DM_NBUF, buf = q & 1;
        if (tid == 0) produce(q + 2);
        const int kind = tile_kind(t);
        cp_async_wait<1>();  // this thread's fragments of tile t have landed (tile t+1 may be in flight)
        if (kind) {
            const double2* st = reinterpret_cast<const double2*>(sP + buf * DM_STAGE) + tid;
#pragma unroll
            for (int bi = 0; bi < 4; bi++)
#pragma unroll
                for (int bj = 0; bj < 2; bj++) {
                    const double2 v = st[(bi * 2 + bj) * DM_THREADS];
                    f[bi][bj][0] = v.x;
                    f[bi][bj][1] = v.y;
                }
        }
        // every warp waits (also one with nothing to do): a warp may not run ahead of the panel ring
        mbar_wait(bar_full + 8 * s, (q / DM_NBUF) & 1);
        const double* pc = sC + s * (DM_K * DM_SC) + lc * DM_SC + wc * 16 + lr;
        if (kind) dmma_ksteps(f, pr, pc, 0, 1);   // consumes the staged values: the slot may be refilled
        fetch_tile(t + 2, tile_kind(t + 2), buf);
        if (kind) {
            dmma_ksteps(f, pr, pc, 1, nks);
            if (!(dbg & 2)) {
                const int jw = t * DM_TN + wc * 16;
#pragma unroll
                for (int bi = 0; bi < 4; bi++) {
                    const int i = iw + bi * 8 + lr;
#pragma unroll
                    for (int bj = 0; bj < 2; bj++) {
                        const int j = jw + bj * 8 + 2 * lc;
                        // pairs straddling the diagonal (j + 1 == i) rewrite one unauthoritative lower element
                        if (kind == 1 || (i < n && j < n && j + 1 >= i)) {
                            double v0 = f[bi][bj][0], v1 = f[bi][bj][1];
                            if (kind == 2 && diag_eps != 0.0) {
                                if (i == j) v0 += diag_eps;
                                if (i == j + 1) v1 += diag_eps;
                            }
                            __stcs(reinterpret_cast<double2*>(prow[bi] + jw + bj * 8 + dshift), make_double2(v0, v1));
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * s);  // this warp is done reading panel buffer s
    }
    cp_async_wait<0>();
}

// Ar / Ac: pre-tiled panel buffers owned by the handle (dmma_panel_doubles() each).
size_t dmma_panel_doubles(int n_cap) {
    const size_t ntr = ((size_t)n_cap + DM_TM - 1) / DM_TM;
    return ntr * DM_K * DM_SR + (DM_TM / DM_TN) * ntr * DM_K * DM_SC;  // Ar then Ac (4 column tiles per strip)
}

int g_dmma_dbg = 0;  // development knob of tools/dmma_bench.cu (1: skip P loads, 2: skip P stores)

int launch_cov_update_dmma(double* P, size_t ld, int n, const double* A, size_t lda, int r, Shard sh,
                           double* panels, int n_cap, int chunk, cudaStream_t stream, double* Pdst = nullptr,
                           double diag_eps = 0.0) {
    const int rp = (r + 3) / 4 * 4;
    const int ntr = (n + DM_TM - 1) / DM_TM;
    const int nbc = (n + DM_TN - 1) / DM_TN;
    const size_t ntr_cap = ((size_t)n_cap + DM_TM - 1) / DM_TM;
    double* Ar = panels;
    double* Ac = panels + ntr_cap * DM_K * DM_SR;
    if (chunk <= 0) chunk = 32;
    count_launch();
    k_dmma_panels<<<dim3((ntr * DM_TM + 255) / 256, rp), 256, 0, stream>>>(A, lda, n, r, ntr * DM_TM, Ar, Ac);
    CSLAM_CUDA(cudaGetLastError());
    int strips = 0;
    for (int tr = sh.rank; tr < ntr; tr += sh.world) strips++;
    if (strips == 0) return CSLAM_OK;
    const dim3 grid((nbc + chunk - 1) / chunk, strips);
    const ptrdiff_t dshift = Pdst != nullptr ? Pdst - P : 0;
    count_launch();
    int dev = 0;
    CSLAM_CUDA(cudaGetDevice(&dev));
#define DM_LAUNCH(FULL)                                                                                        \
    do {                                                                                                       \
        static bool attr_set[64] = {}; /* per device: the attribute is set once, not per call */               \
        if (dev < 0 || dev >= 64 || !attr_set[dev]) {                                                          \
            CSLAM_CUDA(cudaFuncSetAttribute(k_cov_update_dmma<FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                            DM_SMEM));                                                         \
            if (dev >= 0 && dev < 64) attr_set[dev] = true;                                                    \
        }                                                                                                      \
        k_cov_update_dmma<FULL><<<grid, DM_THREADS, DM_SMEM, stream>>>(P, ld, n, Ar, Ac, rp, nbc, chunk, sh,   \
                                                                       g_dmma_dbg, dshift, diag_eps);          \
    } while (0)
    if (rp == DM_K) DM_LAUNCH(true); else DM_LAUNCH(false);
#undef DM_LAUNCH
    CSLAM_CUDA(cudaGetLastError());
    return CSLAM_OK;
}

}  // namespace cslam
