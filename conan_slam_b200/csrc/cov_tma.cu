// cov_tma.cu — the streaming covariance pass of a GROUP of sequential updates,
//     P <- P - sum_q a_q a_q^T      (slam.h:260 for the g observations of one scan, EKF.cpp:457-479;
//                                    panel rows 2q, 2q+1 of A are W1 of observation q)
// as a persistent TMA + FP64-tensor-core kernel for sm_100a.
//
// Why: the FMA form of this pass (k_cov_update_multi, cov_update.cuh) is issue-bound, not memory-bound
// (ncu, round 1: 99 warp instructions per 512 bytes of P at g = 4 — address arithmetic, predicates,
// 128-bit LDG/STG and 8g non-fused FP64 operations; issue slots 48 % busy, 0.87 of HBM at g = 4, 0.72 at
// g = 8).  Here the covariance never passes through the load/store pipeline of the SM cores at all:
//   * P tiles travel HBM -> shared memory -> HBM by tensor-map TMA (cp.async.bulk.tensor.2d, SASS
//     UTMALDG / UTMASTG): one elected thread issues 4 KB boxes, a 5-stage ring of 32 KB sub-tiles
//     (32 rows x 128 columns) keeps >= 64 KB of loads in flight per SM independent of the compute;
//   * the rank-2g term is ONE contraction on the FP64 tensor cores (mma.sync m8n8k4 -> DMMA.8x8x4):
//     2g/4 DMMAs per 8x8 block instead of 8g FP64 instructions per 16-byte pair x 32 lanes;
//   * the accumulator fragments are read from / written to the swizzled tile with conflict-free
//     128-bit LDS/STS; the row / column panels of the tile arrive by 1-D TMA bulk copies.
// Per 32 KB sub-tile a warp issues ~75 instructions (g = 8) where the FMA kernel issued ~1400.
//
// Arithmetic: DMMA accumulates the 2g products with fused multiply-adds in k order on top of P(i,j);
// the FMA kernel subtracts g separately rounded terms.  Both are correctly rounded evaluations of the
// same expression; they differ in the last bits (~1e-16 relative), far inside the 1e-9 parity budget
// (DESIGN.md §2).  Association decisions are taken by the gate kernel, never here.
//
// Layout facts the kernel relies on:
//   * TMA box = 16 columns x 32 rows of doubles (128 B inner dimension, SWIZZLE_128B): the 16-byte chunk c
//     of row r lands at chunk (c ^ (r & 7)) of that row's 128 bytes; stage and box bases are 1024-aligned.
//   * DMMA fragment: lane l holds C[g][2t], C[g][2t+1] with g = l / 4, t = l % 4.  The 8 DMMA rows of a block
//     are mapped to tile rows rho(g) = 4 (g & 1) + (g >> 1) so that the 8 lanes of a quarter-warp (two g) touch
//     rows that differ in bit 2 -> after the swizzle their 16-byte chunks cover all 32 banks exactly once.
//     The row panel is read with the same permutation (any row permutation is legal: the update is
//     elementwise in (i, j) once A and C agree on it).
//   * Row-panel row k starts at 132 k - 2 (k & 1) doubles: with the rho permutation the 16 lanes of a
//     half-warp then hit 16 distinct 8-byte banks; column-panel rows have stride 132 (natural g order).
//   * Everything inside a 128 x 128 tile of the upper triangle is updated, also the (unauthoritative)
//     lower half of diagonal tiles and rows / columns between n and the capacity: nothing reads them
//     before augmentation overwrites them (k_augment writes every upper entry of a new column).
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "ptx_async.cuh"

namespace cslam {

using namespace ptx;

constexpr int TM_T = 128;                          // tile edge (= shard block rows)
constexpr int TM_SUB = 32;                         // rows per pipeline stage
constexpr int TM_BOXC = 16;                        // columns per TMA box (128 B, the swizzle span)
constexpr int TM_BOX_BYTES = TM_BOXC * TM_SUB * 8; // 4 KB
constexpr int TM_STAGE_BYTES = TM_T * TM_SUB * 8;  // 32 KB = 8 boxes
constexpr int TM_PITCH = 132;                      // panel row pitch (doubles)
constexpr int TM_CONSUMERS = 8;                    // consumer warps, one 16-column box each
constexpr int TM_THREADS = (TM_CONSUMERS + 1) * 32;

template <int KS, int S>
struct TmaSmem {
    static constexpr int RP = 4 * KS;
    static constexpr int ring_bytes = S * TM_STAGE_BYTES;
    static constexpr int panel_doubles = RP * TM_PITCH;            // one panel (row or column) of one slot
    static constexpr int panels_bytes = 2 * 2 * panel_doubles * 8;  // 2 slots x (row, column)
    static constexpr int bar_count = 2 * S + 4;                     // full[S], free[S], pfull[2], pempty[2] (even: an int4 follows)
    static constexpr int info_bytes = 2 * 16;                       // per panel slot: {j0, local row 0, diagonal tile?}
    static constexpr int total = ring_bytes + panels_bytes + bar_count * 8 + info_bytes;
};

__device__ __forceinline__ int tm_row_base(int k) { return TM_PITCH * k - 2 * (k & 1); }

// KS = padded rank / 4 (1..4), S = ring depth.  r = actual number of panel rows (<= 4 KS).
// grid = min(#SMs, tiles); CTA b handles tiles b, b + grid, ... of this rank's row-major triangle.
// tmSrc / tmDst: tensor maps of the covariance the pass reads / writes (the same map for an in-place pass, the
// two halves of a ping-pong pair otherwise).  diag_eps: added once to every diagonal element (the sum of the
// FLT_MIN terms of the heading updates in this group, slam.h:719).
// Roles: warp 8 lane 0 = producer (panel bulk copies + tile loads); warps 0..7 = consumers, each owns the 16-column
// box `warp` of every stage: it waits for the stage, updates its box in place in shared memory, and its lane 0
// stores the box back by TMA itself — eight independent store pipelines, no CTA-wide hand-over.
template <int KS, int S>
__global__ void __launch_bounds__(TM_THREADS, 1) k_cov_update_tma(const __grid_constant__ CUtensorMap tmSrc,
                                                                 const __grid_constant__ CUtensorMap tmDst,
                                                                 const double* __restrict__ A, size_t lda, int r,
                                                                 int nt, long long tiles, Shard sh,
                                                                 const int* __restrict__ live, int nlive,
                                                                 double diag_eps, int dbg, int boxr) {
    using L = TmaSmem<KS, S>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* ring = smem_raw;
    double* panels = reinterpret_cast<double*>(smem_raw + L::ring_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L::ring_bytes + L::panels_bytes);
    int4* tinfo = reinterpret_cast<int4*>(smem_raw + L::ring_bytes + L::panels_bytes + L::bar_count * 8);

    // fused scan: if no observation of the group passed the gate every panel is zero — skip the pass
    if (live != nullptr) {
        int any = 0;
        for (int q = 0; q < nlive; q++) any |= live[q];
        if (!any) return;
    }

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t bar_full = bar0, bar_free = bar0 + 8 * S, bar_pfull = bar0 + 16 * S, bar_pempty = bar0 + 16 * S + 16;
    auto rowp = [&](int slot) { return panels + (size_t)slot * 2 * L::panel_doubles; };
    auto colp = [&](int slot) { return panels + (size_t)slot * 2 * L::panel_doubles + L::panel_doubles; };

    if (tid == 0) {
        if ((smem_u32(smem_raw) & 1023u) != 0) __trap();  // SWIZZLE_128B needs 1024-byte aligned stages
#pragma unroll
        for (int s = 0; s < S; s++) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_free + 8 * s, TM_CONSUMERS);
        }
#pragma unroll
        for (int s = 0; s < 2; s++) {
            mbar_init(bar_pfull + 8 * s, 1);
            mbar_init(bar_pempty + 8 * s, TM_CONSUMERS);
        }
        mbar_fence_init();
    }
    // Panels start out zero: rows r..RP-1 (rank padding) stay zero for the whole kernel — the bulk copies only
    // ever write rows < r.  The proxy fence orders these generic-proxy writes before the async-proxy copies.
    for (int idx = tid; idx < 4 * L::panel_doubles / 2; idx += TM_THREADS)
        reinterpret_cast<double2*>(panels)[idx] = make_double2(0.0, 0.0);
    fence_proxy_async();
    __syncthreads();

    if (warp == TM_CONSUMERS) {
        // ------------------------------------------------------------------ producer ----
        if (lane != 0) return;
        tmap_prefetch(&tmSrc);
        const uint64_t pol_stream = policy_evict_first();  // P: touched once per pass
        const uint64_t pol_keep = policy_evict_last();     // panels: re-read by every tile of a strip / column
        const uint32_t ring_u32 = smem_u32(ring);
        long long sub = 0;
        int st = 0;
        uint32_t free_phase = 0;  // bit s = parity to wait for on free[s]
        int lt = 0;
        for (long long t = blockIdx.x; t < tiles; t += gridDim.x, lt++) {
            int tr, tc;
            shard_tile(t, nt, sh, tr, tc);
            const int slot = lt & 1;
            if (lt >= 2) mbar_wait(bar_pempty + 8 * slot, ((lt >> 1) - 1) & 1);
            const int i0 = tr * TM_T, j0 = tc * TM_T;
            const int lrow0 = (int)shard_lrow(sh, i0);
            tinfo[slot] = make_int4(j0, lrow0, tr == tc ? 1 : 0, 0);  // published by the arrive below (release)
            const int ilen = min(TM_T, (int)lda - i0), jlen = min(TM_T, (int)lda - j0);  // doubles, > 0, even
            mbar_expect_tx(bar_pfull + 8 * slot, (uint32_t)(r * (ilen + jlen) * 8));
            {
                const uint32_t rp_u32 = smem_u32(rowp(slot)), cp_u32 = smem_u32(colp(slot));
                for (int k = 0; k < r; k++) {
                    bulk_g2s_hint(rp_u32 + 8 * tm_row_base(k), A + (size_t)k * lda + i0, (uint32_t)ilen * 8,
                                  bar_pfull + 8 * slot, pol_keep);
                    bulk_g2s_hint(cp_u32 + 8 * TM_PITCH * k, A + (size_t)k * lda + j0, (uint32_t)jlen * 8,
                                  bar_pfull + 8 * slot, pol_keep);
                }
            }
#pragma unroll 1
            for (int s = 0; s < TM_T / TM_SUB; s++) {
                if (sub >= S) {  // every consumer warp's store of the sub-tile that used this stage has left shared memory
                    mbar_wait(bar_free + 8 * st, (free_phase >> st) & 1u);
                    free_phase ^= 1u << st;
                }
                if (dbg & 4) {  // development ablation: no covariance loads
                    mbar_arrive(bar_full + 8 * st);
                } else {
                    mbar_expect_tx(bar_full + 8 * st, TM_STAGE_BYTES);
                    // row-group major: the 8 boxes that cover the same rows (1 KB contiguous per row) are requested
                    // back to back
                    for (int rg = 0; rg < TM_SUB; rg += boxr)
#pragma unroll
                        for (int b = 0; b < 8; b++)
                            tma_load_2d(ring_u32 + st * TM_STAGE_BYTES + b * TM_BOX_BYTES + rg * 128, &tmSrc,
                                        j0 + TM_BOXC * b, lrow0 + TM_SUB * s + rg, bar_full + 8 * st, pol_stream);
                }
                st = st + 1 == S ? 0 : st + 1;
                sub++;
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers ----
    const int g = lane >> 2, t = lane & 3;
    const int rho = ((g & 1) << 2) | (g >> 1);
    // fragment offsets inside this warp's box: row rho, 16-byte chunk (4 cb + t) ^ rho
    const uint32_t off0 = (uint32_t)(rho * 128 + ((t ^ rho) << 4));
    const uint32_t off1 = off0 ^ 64u;
    const long long my_tiles = (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const uint64_t pol_stream = policy_evict_first();
    if (lane == 0) tmap_prefetch(&tmDst);
    int st = 0, prev_st = -1;
    uint32_t full_phase = 0;
    for (long long lt = 0; lt < my_tiles; lt++) {
        const int slot = (int)(lt & 1);
        mbar_wait(bar_pfull + 8 * slot, (uint32_t)((lt >> 1) & 1));
        const int4 ti = tinfo[slot];
        // negated column-panel fragments of this warp's 16 columns: constant over the tile
        double nb[KS][2];
        {
            const double* cp = colp(slot) + TM_BOXC * warp + g;
#pragma unroll
            for (int ks = 0; ks < KS; ks++)
#pragma unroll
                for (int cb = 0; cb < 2; cb++) nb[ks][cb] = -cp[TM_PITCH * (4 * ks + t) + 8 * cb];
        }
        const double* rp = rowp(slot) + rho;
#pragma unroll 1
        for (int s = 0; s < TM_T / TM_SUB; s++) {
            mbar_wait(bar_full + 8 * st, (full_phase >> st) & 1u);
            full_phase ^= 1u << st;
            unsigned char* box = ring + st * TM_STAGE_BYTES + warp * TM_BOX_BYTES;
            double2 acc[4][2];
#pragma unroll
            for (int rb = 0; rb < 4; rb++) {
                acc[rb][0] = *reinterpret_cast<const double2*>(box + rb * 1024 + off0);
                acc[rb][1] = *reinterpret_cast<const double2*>(box + rb * 1024 + off1);
            }
#pragma unroll
            for (int ks = 0; ks < KS; ks++) {
                double a[4];
#pragma unroll
                for (int rb = 0; rb < 4; rb++) a[rb] = rp[tm_row_base(4 * ks + t) + TM_SUB * s + 8 * rb];
#pragma unroll
                for (int rb = 0; rb < 4; rb++) {
                    dmma884(acc[rb][0].x, acc[rb][0].y, a[rb], nb[ks][0]);
                    dmma884(acc[rb][1].x, acc[rb][1].y, a[rb], nb[ks][1]);
                }
            }
            if (ti.z && diag_eps != 0.0) {  // diagonal tile: slam.h:719 on the diagonal elements
#pragma unroll
                for (int rb = 0; rb < 4; rb++) {
                    const int row = TM_SUB * s + 8 * rb + rho;  // tile-relative
#pragma unroll
                    for (int cb = 0; cb < 2; cb++) {
                        const int col = TM_BOXC * warp + 8 * cb + 2 * t;
                        if (col == row) acc[rb][cb].x += diag_eps;
                        if (col + 1 == row) acc[rb][cb].y += diag_eps;
                    }
                }
            }
#pragma unroll
            for (int rb = 0; rb < 4; rb++) {
                *reinterpret_cast<double2*>(box + rb * 1024 + off0) = acc[rb][0];
                *reinterpret_cast<double2*>(box + rb * 1024 + off1) = acc[rb][1];
            }
            fence_proxy_async();  // the TMA store (async proxy) must see these generic-proxy writes
            __syncwarp();
            if (lane == 0) {
                if (!(dbg & 2))
                    for (int rg = 0; rg < TM_SUB; rg += boxr)
                        tma_store_2d(&tmDst, ti.x + TM_BOXC * warp, ti.y + TM_SUB * s + rg, smem_u32(box) + rg * 128,
                                     pol_stream);
                bulk_commit();
                if (prev_st >= 0) {  // the previous store of this warp has left shared memory: its stage is free
                    bulk_wait_read<1>();
                    mbar_arrive(bar_free + 8 * prev_st);
                }
            }
            prev_st = st;
            st = st + 1 == S ? 0 : st + 1;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_pempty + 8 * slot);
    }
    if (lane == 0) bulk_wait<0>();  // all writes performed before the kernel ends
}

// ---------------------------------------------------------------------------------------------------
// Dense-row variant (the default): a stage is 32 full rows of the tile (32 x 1 KB, no swizzle), ONE tensor-map
// load and ONE tensor-map store per stage, both 1 KB contiguous per covariance row.
// Roles: warps 0..7 consumers (each the 16-column slice `warp` of every stage), warp 8 lane 0 loads (panels +
// tiles), warps 9..9+TMD_STORERS-1 lane 0 store: store thread j owns the sub-tiles u = j (mod TMD_STORERS), issues
// the store when the consumers are done and waits for ITS OWN store to have left shared memory before it frees
// the stage — frees are prompt and TMD_STORERS stores are in flight (measured: one store thread that frees a
// stage only when it issues the next store costs a whole stage of the ring).
// Shared-memory budget (227 KB): the ring wants every byte — B200 needs ~50 KB of loads AND ~50 KB of stores in
// flight per SM to saturate HBM — so the column panel is single-buffered: the consumers copy their fragments
// into registers at the start of a tile and hand the buffer back at once; the row panel (re-read by every
// sub-tile) is double-buffered.
// Bank conflicts of the fragment accesses (row pitch 1 KB: the two DMMA rows of a quarter-warp would hit the
// same banks) are avoided by letting the lanes of odd DMMA rows touch the warp's second 8-column block first
// and swapping the two registers afterwards.
constexpr int TMD_STORERS = 2;
constexpr int TMD_THREADS = (TM_CONSUMERS + 1 + TMD_STORERS) * 32;
template <int KS, int S, int SUB>
struct TmdSmem {
    static constexpr int RP = 4 * KS;
    static constexpr int stage_bytes = SUB * TM_T * 8;  // SUB rows x 1 KB
    static constexpr int ring_bytes = S * stage_bytes;
    static constexpr int panel_doubles = RP * TM_PITCH;
    static constexpr int panels_bytes = 3 * panel_doubles * 8;  // row panel x 2 slots, column panel x 1
    static constexpr int bar_count = (3 * S + 6 + 1) / 2 * 2;    // full, done, free [S]; pfull[2], pempty[2], cfull, cempty (even: int4 behind)
    static constexpr int total = ring_bytes + panels_bytes + bar_count * 8 + 2 * 16;
};
// SUB = rows per pipeline stage (16 or 32): the ring holds S x SUB KB.
template <int KS, int S, int SUB>
__global__ void __launch_bounds__(TMD_THREADS, 1) k_cov_update_tma_dense(const __grid_constant__ CUtensorMap tmSrc,
                                                                        const __grid_constant__ CUtensorMap tmDst,
                                                                        const double* __restrict__ A, size_t lda,
                                                                        int r, int nt, long long tiles, Shard sh,
                                                                        const int* __restrict__ live, int nlive,
                                                                        double diag_eps, int dbg, double* Pdst,
                                                                        size_t ld, int rows_cap) {
    // Pdst != nullptr: "direct store" mode — the consumers write their accumulator fragments straight to global
    // memory (streaming 16-byte stores) instead of handing the stage to a TMA store; a stage is then free as soon
    // as it has been read into registers.
    using L = TmdSmem<KS, S, SUB>;
    constexpr int RB = SUB / 8;               // 8-row DMMA blocks per stage and warp
    constexpr int STAGE = L::stage_bytes;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* ring = smem_raw;
    double* panels = reinterpret_cast<double*>(smem_raw + L::ring_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L::ring_bytes + L::panels_bytes);
    int4* tinfo = reinterpret_cast<int4*>(smem_raw + L::ring_bytes + L::panels_bytes + L::bar_count * 8);
    if (live != nullptr) {
        int any = 0;
        for (int q = 0; q < nlive; q++) any |= live[q];
        if (!any) return;
    }
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t bar_full = bar0, bar_done = bar0 + 8 * S, bar_free = bar0 + 16 * S, bar_pfull = bar0 + 24 * S,
                   bar_pempty = bar_pfull + 16, bar_cfull = bar_pfull + 32, bar_cempty = bar_pfull + 40;
    const uint32_t skew = (uint32_t)dbg & 0x40000000u;  // zero at run time, unknown to the assembler (mbar_arrive_after)
    auto rowp = [&](int slot) { return panels + (size_t)slot * L::panel_doubles; };
    double* colp = panels + 2 * (size_t)L::panel_doubles;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < S; s++) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_free + 8 * s, Pdst != nullptr ? TM_CONSUMERS : 1);
            mbar_init(bar_done + 8 * s, TM_CONSUMERS);
        }
#pragma unroll
        for (int s = 0; s < 2; s++) {
            mbar_init(bar_pfull + 8 * s, 1);
            mbar_init(bar_pempty + 8 * s, TM_CONSUMERS);
        }
        mbar_init(bar_cfull, 1);
        mbar_init(bar_cempty, TM_CONSUMERS);
        mbar_fence_init();
    }
    // panels start out zero: rows r..RP-1 (rank padding) stay zero, the bulk copies only write rows < r
    for (int idx = tid; idx < 3 * L::panel_doubles / 2; idx += TMD_THREADS)
        reinterpret_cast<double2*>(panels)[idx] = make_double2(0.0, 0.0);
    fence_proxy_async();
    __syncthreads();

    if (warp == TM_CONSUMERS) {  // ------------------------------------------------ loads ----
        if (lane != 0) return;
        tmap_prefetch(&tmSrc);
        const uint64_t pol_stream = policy_evict_first(), pol_keep = policy_evict_last();
        const uint32_t ring_u32 = smem_u32(ring);
        long long sub = 0;
        int st = 0, lt = 0;
        uint32_t free_phase = 0;
        for (long long t = blockIdx.x; t < tiles; t += gridDim.x, lt++) {
            int tr, tc;
            shard_tile(t, nt, sh, tr, tc);
            const int slot = lt & 1;
            const int i0 = tr * TM_T, j0 = tc * TM_T;
            const int lrow0 = (int)shard_lrow(sh, i0);
            const int ilen = min(TM_T, (int)lda - i0), jlen = min(TM_T, (int)lda - j0);  // doubles, > 0, even
            if (lt >= 2) mbar_wait(bar_pempty + 8 * slot, ((lt >> 1) - 1) & 1);
            tinfo[slot] = make_int4(j0, lrow0, tr == tc ? 1 : 0, 0);  // published by the arrive below (release)
            mbar_expect_tx(bar_pfull + 8 * slot, (uint32_t)(r * ilen * 8));
            const uint32_t rp_u32 = smem_u32(rowp(slot)), cp_u32 = smem_u32(colp);
            for (int k = 0; k < r; k++)
                bulk_g2s_hint(rp_u32 + 8 * TM_PITCH * k, A + (size_t)k * lda + i0, (uint32_t)ilen * 8, bar_pfull + 8 * slot,
                              pol_keep);
            if (lt >= 1) mbar_wait(bar_cempty, (uint32_t)((lt - 1) & 1));  // fragments of the previous tile are in registers
            mbar_expect_tx(bar_cfull, (uint32_t)(r * jlen * 8));
            for (int k = 0; k < r; k++)
                bulk_g2s_hint(cp_u32 + 8 * TM_PITCH * k, A + (size_t)k * lda + j0, (uint32_t)jlen * 8, bar_cfull, pol_keep);
#pragma unroll 1
            for (int s = 0; s < TM_T / SUB; s++) {
                if (sub >= S) {
                    mbar_wait(bar_free + 8 * st, (free_phase >> st) & 1u);
                    free_phase ^= 1u << st;
                }
                if (dbg & 4) {  // development ablation: no covariance loads
                    mbar_arrive(bar_full + 8 * st);
                } else {
                    mbar_expect_tx(bar_full + 8 * st, STAGE);
                    tma_load_2d(ring_u32 + st * STAGE, &tmSrc, j0, lrow0 + SUB * s, bar_full + 8 * st, pol_stream);
                }
                st = st + 1 == S ? 0 : st + 1;
                sub++;
            }
        }
        return;
    }
    if (warp > TM_CONSUMERS) {  // ------------------------------------------------- stores ----
        if (lane != 0 || Pdst != nullptr) return;
        const int me = warp - TM_CONSUMERS - 1;
        tmap_prefetch(&tmDst);
        const uint64_t pol_stream = policy_evict_first();
        const uint32_t ring_u32 = smem_u32(ring);
        int st = 0;
        long long sub = 0;
        uint32_t done_phase = 0;
        for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
            int tr, tc;
            shard_tile(t, nt, sh, tr, tc);
            const int j0 = tc * TM_T, lrow0 = (int)shard_lrow(sh, tr * TM_T);
#pragma unroll 1
            for (int s = 0; s < TM_T / SUB; s++, sub++) {
                if ((int)(sub % TMD_STORERS) == me) {
                    mbar_wait(bar_done + 8 * st, (done_phase >> st) & 1u);
                    if (!(dbg & 2)) tma_store_2d(&tmDst, j0, lrow0 + SUB * s, ring_u32 + st * STAGE, pol_stream);
                    bulk_commit();
                    bulk_wait_read<0>();  // this store has left shared memory: the stage is free
                    mbar_arrive(bar_free + 8 * st);
                }
                done_phase ^= 1u << st;  // every storer tracks the phase of every stage
                st = st + 1 == S ? 0 : st + 1;
            }
        }
        bulk_wait<0>();  // all writes performed before the kernel ends
        return;
    }
    // ---------------------------------------------------------------------------- consumers ----
    const int g = lane >> 2, t = lane & 3, odd = g & 1;
    // lanes of odd DMMA rows read the warp's second 8-column block with the first instruction
    const uint32_t offA = (uint32_t)(g * 1024 + (TM_BOXC * warp + 8 * odd + 2 * t) * 8);
    const uint32_t offB = (uint32_t)(g * 1024 + (TM_BOXC * warp + 8 * (odd ^ 1) + 2 * t) * 8);
    const long long my_tiles = (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    int st = 0;
    uint32_t full_phase = 0;
    for (long long lt = 0; lt < my_tiles; lt++) {
        const int slot = (int)(lt & 1);
        // negated column-panel fragments of this warp's 16 columns: constant over the tile, kept in registers;
        // the buffer goes straight back to the producer
        mbar_wait(bar_cfull, (uint32_t)(lt & 1));
        double nb[KS][2];
        {
            const double* cp = colp + TM_BOXC * warp + g;
#pragma unroll
            for (int ks = 0; ks < KS; ks++)
#pragma unroll
                for (int cb = 0; cb < 2; cb++) nb[ks][cb] = -cp[TM_PITCH * (4 * ks + t) + 8 * cb];
        }
        {  // hand the buffer back once these loads have landed (see the note at the stage release below)
            uint32_t dep = 0;
#pragma unroll
            for (int ks = 0; ks < KS; ks++) dep |= hi32(nb[ks][0]) | hi32(nb[ks][1]);
            __syncwarp();
            if (lane == 0) mbar_arrive_after(bar_cempty, skew, dep);
        }
        mbar_wait(bar_pfull + 8 * slot, (uint32_t)((lt >> 1) & 1));
        const int4 ti = tinfo[slot];
        const double* rp = rowp(slot) + g;
#pragma unroll 1
        for (int s = 0; s < TM_T / SUB; s++) {
            mbar_wait(bar_full + 8 * st, (full_phase >> st) & 1u);
            full_phase ^= 1u << st;
            unsigned char* stage = ring + st * STAGE;
            double2 acc[RB][2];
            uint32_t dep = 0;
#pragma unroll
            for (int rb = 0; rb < RB; rb++) {
                const double2 va = *reinterpret_cast<const double2*>(stage + rb * 8192 + offA);
                const double2 vb = *reinterpret_cast<const double2*>(stage + rb * 8192 + offB);
                acc[rb][0] = odd ? vb : va;
                acc[rb][1] = odd ? va : vb;
                dep |= hi32(va.x) | hi32(vb.x);
            }
            // Direct-store mode hands the stage back as soon as it has been read — i.e. once the loads above have
            // LANDED: an arrive merely placed behind them is issued while they may still sit in the load/store
            // queue (behind this warp's streaming global stores of the previous sub-tile, which stall when HBM is
            // saturated); the producer then refills the stage and the queued loads read the NEXT sub-tile's first
            // rows.  Seen as 4-16 wrong elements in the first rows of a sub-tile once in ~10 passes with rings of
            // 3-5 stages (tools/cov_tma_bench.cu, TMA_DET=<rounds>); mbar_arrive_after carries a register
            // dependency on every load.
            if (Pdst != nullptr) {
                __syncwarp();
                if (lane == 0) mbar_arrive_after(bar_free + 8 * st, skew, dep);
            }
#pragma unroll
            for (int ks = 0; ks < KS; ks++) {
                double a[RB];
#pragma unroll
                for (int rb = 0; rb < RB; rb++) a[rb] = rp[TM_PITCH * (4 * ks + t) + SUB * s + 8 * rb];
#pragma unroll
                for (int rb = 0; rb < RB; rb++) {
                    dmma884(acc[rb][0].x, acc[rb][0].y, a[rb], nb[ks][0]);
                    dmma884(acc[rb][1].x, acc[rb][1].y, a[rb], nb[ks][1]);
                }
            }
            if (ti.z && diag_eps != 0.0) {  // diagonal tile: slam.h:719 on the diagonal elements
#pragma unroll
                for (int rb = 0; rb < RB; rb++) {
                    const int row = SUB * s + 8 * rb + g;
#pragma unroll
                    for (int cb = 0; cb < 2; cb++) {
                        const int col = TM_BOXC * warp + 8 * cb + 2 * t;
                        if (col == row) acc[rb][cb].x += diag_eps;
                        if (col + 1 == row) acc[rb][cb].y += diag_eps;
                    }
                }
            }
            if (Pdst != nullptr) {
                // lane (g, t) owns columns 2t, 2t+1 of row g of each 8x8 block: 64 contiguous bytes per row and block
                const int col0 = ti.x + TM_BOXC * warp + 2 * t;
                double* prow = Pdst + (size_t)(ti.y + SUB * s + g) * ld + col0;
#pragma unroll
                for (int rb = 0; rb < RB; rb++) {
                    if (ti.y + SUB * s + 8 * rb + g < rows_cap && !(dbg & 2)) {
                        if (col0 < (int)ld) __stcs(reinterpret_cast<double2*>(prow + (size_t)8 * rb * ld), acc[rb][0]);
                        if (col0 + 8 < (int)ld) __stcs(reinterpret_cast<double2*>(prow + (size_t)8 * rb * ld + 8), acc[rb][1]);
                    }
                }
            } else {
#pragma unroll
                for (int rb = 0; rb < RB; rb++) {
                    *reinterpret_cast<double2*>(stage + rb * 8192 + offA) = odd ? acc[rb][1] : acc[rb][0];
                    *reinterpret_cast<double2*>(stage + rb * 8192 + offB) = odd ? acc[rb][0] : acc[rb][1];
                }
                fence_proxy_async();  // the TMA store (async proxy) must see these generic-proxy writes
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_done + 8 * st);
            }
            st = st + 1 == S ? 0 : st + 1;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_pempty + 8 * slot);
    }
}

// ---------------------------------------------------------------------------------------------------
// Joint (batch) update, rank r <= 64 (EKF.cpp:93-129 -> slam.h:260 with r = 2m, m <= 32): the SAME dense-row
// pipeline, now bound by the FP64 tensor cores instead of HBM (8 flop per byte at r = 64).  What changes:
//   * the column-panel fragments of a tile (r/4 x 2 doubles per lane, up to 64 registers) live in registers for
//     the whole tile; the shared-memory copy is single-buffered and handed back as soon as they are loaded;
//   * the row panel (r x 128 doubles = 66 KB at r = 64) is single-buffered too: a CTA walks CHUNKS of consecutive
//     tiles of the row-major triangle, so the panel is reloaded only when the 128-row strip changes;
//   * the A fragments of k-step s+1 are fetched from shared memory while the DMMAs of k-step s issue;
//   * results leave the SM by streaming 16-byte stores straight from the accumulator fragments (direct-store
//     mode), so a stage is free the moment it has been read into registers and a 2-deep ring suffices.
// P travels HBM -> shared memory by tensor-map TMA (UTMALDG); no LDG / LDGSTS / address arithmetic per element
// competes with the DMMA issue slots (round 1's kernel staged P with per-thread cp.async: 0.81 of the DMMA peak,
// its own no-load/no-store ceiling was 0.90).
constexpr int TMJ_THREADS = (TM_CONSUMERS + 1) * 32;
constexpr int TMJ_KS = 16;  // k-steps at full rank (64 / 4)
template <int S, int SUB>
struct TmjSmem {
    static constexpr int stage_bytes = SUB * TM_T * 8;
    static constexpr int ring_bytes = S * stage_bytes;
    static constexpr int panel_doubles = 4 * TMJ_KS * TM_PITCH;
    static constexpr int panels_bytes = 2 * panel_doubles * 8;  // row panel, column panel (one slot each)
    static constexpr int bar_count = 2 * S + 4;                 // full[S], free[S], rfull, rempty, cfull, cempty
    static constexpr int total = ring_bytes + panels_bytes + bar_count * 8 + 32;
};
template <int S, int SUB>
__global__ void __launch_bounds__(TMJ_THREADS, 1) k_cov_update_tma_joint(const __grid_constant__ CUtensorMap tmSrc,
                                                                        const double* __restrict__ A, size_t lda,
                                                                        int r, int nt, long long tiles, int chunk,
                                                                        Shard sh, double* __restrict__ Pdst, size_t ld,
                                                                        int rows_cap) {
    using L = TmjSmem<S, SUB>;
    constexpr int RB = SUB / 8;
    constexpr int STAGE = L::stage_bytes;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* ring = smem_raw;
    double* rowp = reinterpret_cast<double*>(smem_raw + L::ring_bytes);
    double* colp = rowp + L::panel_doubles;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L::ring_bytes + L::panels_bytes);
    int4* tinfo = reinterpret_cast<int4*>(smem_raw + L::ring_bytes + L::panels_bytes + L::bar_count * 8);  // [2]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t bar_full = bar0, bar_free = bar0 + 8 * S, bar_rfull = bar0 + 16 * S, bar_rempty = bar_rfull + 8,
                   bar_cfull = bar_rfull + 16, bar_cempty = bar_rfull + 24;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < S; s++) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_free + 8 * s, TM_CONSUMERS);
        }
        mbar_init(bar_rfull, 1);
        mbar_init(bar_rempty, TM_CONSUMERS);
        mbar_init(bar_cfull, 1);
        mbar_init(bar_cempty, TM_CONSUMERS);
        mbar_fence_init();
    }
    for (int idx = tid; idx < 2 * L::panel_doubles / 2; idx += TMJ_THREADS)
        reinterpret_cast<double2*>(rowp)[idx] = make_double2(0.0, 0.0);  // rank padding rows stay zero
    fence_proxy_async();
    __syncthreads();
    // this CTA's tiles: chunks b, b + grid, ... of `chunk` consecutive tiles of the rank's row-major triangle
    const long long nchunks = (tiles + chunk - 1) / chunk;

    if (warp == TM_CONSUMERS) {  // -------------------------------------------------- producer ----
        if (lane != 0) return;
        tmap_prefetch(&tmSrc);
        const uint64_t pol_stream = policy_evict_first(), pol_keep = policy_evict_last();
        const uint32_t ring_u32 = smem_u32(ring), rp_u32 = smem_u32(rowp), cp_u32 = smem_u32(colp);
        long long sub = 0, lt = 0, rgen = 0;
        int st = 0, loaded_tr = -1;
        uint32_t free_phase = 0;
        for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
            const long long t0 = c * chunk, t1 = (t0 + chunk < tiles) ? t0 + chunk : tiles;
            for (long long t = t0; t < t1; t++, lt++) {
                int tr, tc;
                shard_tile(t, nt, sh, tr, tc);
                // is this the last tile that uses the row panel of strip tr?  (look one tile ahead in MY sequence)
                long long tn = t + 1;
                if (tn >= t1) tn = (c + gridDim.x < nchunks) ? (c + gridDim.x) * chunk : -1;
                int ntr = -1, ntc = 0;
                if (tn >= 0) shard_tile(tn, nt, sh, ntr, ntc);
                const int i0 = tr * TM_T, j0 = tc * TM_T;
                const int lrow0 = (int)shard_lrow(sh, i0);
                const int ilen = min(TM_T, (int)lda - i0), jlen = min(TM_T, (int)lda - j0);
                const bool new_row = tr != loaded_tr;
                if (new_row) {
                    if (rgen >= 1) mbar_wait(bar_rempty, (uint32_t)((rgen - 1) & 1));  // consumers left the old strip
                    mbar_expect_tx(bar_rfull, (uint32_t)(r * ilen * 8));
                    for (int k = 0; k < r; k++)
                        bulk_g2s_hint(rp_u32 + 8 * TM_PITCH * k, A + (size_t)k * lda + i0, (uint32_t)ilen * 8, bar_rfull, pol_keep);
                    loaded_tr = tr;
                    rgen++;
                }
                if (lt >= 1) mbar_wait(bar_cempty, (uint32_t)((lt - 1) & 1));  // previous tile's fragments are in registers
                tinfo[lt & 1] = make_int4(j0, lrow0, (new_row ? 1 : 0) | (ntr != tr ? 2 : 0), 0);
                mbar_expect_tx(bar_cfull, (uint32_t)(r * jlen * 8));
                for (int k = 0; k < r; k++)
                    bulk_g2s_hint(cp_u32 + 8 * TM_PITCH * k, A + (size_t)k * lda + j0, (uint32_t)jlen * 8, bar_cfull, pol_keep);
#pragma unroll 1
                for (int s = 0; s < TM_T / SUB; s++) {
                    if (sub >= S) {
                        mbar_wait(bar_free + 8 * st, (free_phase >> st) & 1u);
                        free_phase ^= 1u << st;
                    }
                    mbar_expect_tx(bar_full + 8 * st, STAGE);
                    tma_load_2d(ring_u32 + st * STAGE, &tmSrc, j0, lrow0 + SUB * s, bar_full + 8 * st, pol_stream);
                    st = st + 1 == S ? 0 : st + 1;
                    sub++;
                }
            }
        }
        return;
    }
    // ---------------------------------------------------------------------------- consumers ----
    const int g = lane >> 2, t = lane & 3, odd = g & 1;
    const uint32_t offA = (uint32_t)(g * 1024 + (TM_BOXC * warp + 8 * odd + 2 * t) * 8);
    const uint32_t offB = (uint32_t)(g * 1024 + (TM_BOXC * warp + 8 * (odd ^ 1) + 2 * t) * 8);
    const int nks = (r + 3) / 4;
    long long my_tiles = 0;
    for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const long long t0 = c * chunk;
        my_tiles += ((t0 + chunk < tiles) ? t0 + chunk : tiles) - t0;
    }
    int st = 0;
    uint32_t full_phase = 0;
    long long rgen = 0;
    const double* rp = rowp + g;
    for (long long lt = 0; lt < my_tiles; lt++) {
        mbar_wait(bar_cfull, (uint32_t)(lt & 1));
        const int4 ti = tinfo[lt & 1];
        double nb[TMJ_KS][2];
        {
            const double* cp = colp + TM_BOXC * warp + g;
#pragma unroll
            for (int ks = 0; ks < TMJ_KS; ks++)
#pragma unroll
                for (int cb = 0; cb < 2; cb++) nb[ks][cb] = ks < nks ? -cp[TM_PITCH * (4 * ks + t) + 8 * cb] : 0.0;
        }
        // (column-panel buffer released after the first sub-tile's DMMAs, see k_cov_update_tma_dense)
        if (ti.z & 1) {  // new strip: its row panel
            mbar_wait(bar_rfull, (uint32_t)(rgen & 1));
            rgen++;
        }
#pragma unroll 1
        for (int s = 0; s < TM_T / SUB; s++) {
            mbar_wait(bar_full + 8 * st, (full_phase >> st) & 1u);
            full_phase ^= 1u << st;
            unsigned char* stage = ring + st * STAGE;
            double2 acc[RB][2];
#pragma unroll
            for (int rb = 0; rb < RB; rb++) {
                const double2 va = *reinterpret_cast<const double2*>(stage + rb * 8192 + offA);
                const double2 vb = *reinterpret_cast<const double2*>(stage + rb * 8192 + offB);
                acc[rb][0] = odd ? vb : va;
                acc[rb][1] = odd ? va : vb;
            }
            const int st_read = st;
            st = st + 1 == S ? 0 : st + 1;
            double a0[RB], a1[RB];
#pragma unroll
            for (int rb = 0; rb < RB; rb++) a0[rb] = rp[TM_PITCH * t + SUB * s + 8 * rb];
#pragma unroll
            for (int ks = 0; ks < TMJ_KS; ks += 2) {
                if (ks < nks) {
                    if (ks + 1 < nks) {
#pragma unroll
                        for (int rb = 0; rb < RB; rb++) a1[rb] = rp[TM_PITCH * (4 * (ks + 1) + t) + SUB * s + 8 * rb];
                    }
#pragma unroll
                    for (int rb = 0; rb < RB; rb++) {
                        dmma884(acc[rb][0].x, acc[rb][0].y, a0[rb], nb[ks][0]);
                        dmma884(acc[rb][1].x, acc[rb][1].y, a0[rb], nb[ks][1]);
                    }
                    if (ks + 1 < nks) {
                        if (ks + 2 < nks) {
#pragma unroll
                            for (int rb = 0; rb < RB; rb++) a0[rb] = rp[TM_PITCH * (4 * (ks + 2) + t) + SUB * s + 8 * rb];
                        }
#pragma unroll
                        for (int rb = 0; rb < RB; rb++) {
                            dmma884(acc[rb][0].x, acc[rb][0].y, a1[rb], nb[ks + 1][0]);
                            dmma884(acc[rb][1].x, acc[rb][1].y, a1[rb], nb[ks + 1][1]);
                        }
                    }
                }
            }
            // the stage / column panel go back once the DMMAs have consumed what was loaded from them
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_free + 8 * st_read);
            if (s == 0) {
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_cempty);
            }
            const int col0 = ti.x + TM_BOXC * warp + 2 * t;
            double* prow = Pdst + (size_t)(ti.y + SUB * s + g) * ld + col0;
#pragma unroll
            for (int rb = 0; rb < RB; rb++) {
                if (ti.y + SUB * s + 8 * rb + g < rows_cap) {
                    if (col0 < (int)ld) __stcs(reinterpret_cast<double2*>(prow + (size_t)8 * rb * ld), acc[rb][0]);
                    if (col0 + 8 < (int)ld) __stcs(reinterpret_cast<double2*>(prow + (size_t)8 * rb * ld + 8), acc[rb][1]);
                }
            }
        }
        if (ti.z & 2) {  // last tile of this strip in my sequence: the row panel may be replaced
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_rempty);
        }
    }
}

// ------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int g_tma_sub = getenv("CSLAM_TMA_SUB") ? atoi(getenv("CSLAM_TMA_SUB")) : 32;  // rows per stage of the dense variant (16 | 32)
int g_tma_boxr = 32;  // development knob: rows per TMA box of the swizzled-box variant (8, 16 or 32)
// 1 = dense-row variant (k_cov_update_tma_dense: one 32 KB tensor-map load / store per stage, measured slightly
// faster on a B200), 0 = swizzled 16-column boxes with consumer-issued stores (k_cov_update_tma).
// CSLAM_TMA_DENSE=0/1 overrides; the tensor maps of a handle are made for one variant.
int g_tma_dense = getenv("CSLAM_TMA_DENSE") ? atoi(getenv("CSLAM_TMA_DENSE")) : 1;

// Tensor map over the locally stored rows of P (row-major doubles, `rows` x `ld`): 16 x 32 boxes, 128-byte swizzle.
int make_cov_tensor_map(void* out_map64, double* P, size_t ld, size_t rows) {
    EncodeTiledFn enc = encode_tiled_fn();
    CSLAM_REQUIRE(enc != nullptr, CSLAM_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    CUtensorMap tm;
    const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
    const cuuint32_t box[2] = {g_tma_dense ? (cuuint32_t)TM_T : (cuuint32_t)TM_BOXC, (cuuint32_t)(g_tma_dense ? g_tma_sub : g_tma_boxr)};
    const cuuint32_t estr[2] = {1, 1};
    CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    if (const char* e = getenv("CSLAM_TMA_PROMO")) promo = (CUtensorMapL2promotion)atoi(e);  // development knob (0..3)
    const CUresult rc = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, P, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, g_tma_dense ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B, promo,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed (CUresult %d; ld=%zu rows=%zu)", (int)rc, ld, rows);
        return CSLAM_ERR_CUDA;
    }
    memcpy(out_map64, &tm, sizeof(tm));
    return CSLAM_OK;
}

int g_tma_dbg = 0;  // development knobs of tools/cov_tma_bench.cu (2: no stores, 4: no loads)

struct DirectStore {  // destination array of the direct-store mode
    double* P;
    size_t ld;
    int rows;
};
// 1 (default) = direct-store mode of the dense variant: tensor-map TMA loads, the consumers store their accumulator
// fragments with streaming 16-byte stores; 0 = tensor-map TMA stores through the ring.  Measured on a B200
// (tools/cov_tma_bench.cu, N = 20k): both land at 2.12-2.25 ms per pass; direct stores need only a 2-deep ring
// (64 KB), which leaves room on every SM for the gate / gain kernels that run beside the pass.
int g_tma_direct = getenv("CSLAM_TMA_DIRECT") ? atoi(getenv("CSLAM_TMA_DIRECT")) : 1;


template <int KS, int S>
static int launch_one(const CUtensorMap& tm, const CUtensorMap& tmd, DirectStore ds, double diag_eps, const double* A, size_t lda, int r, int nt, long long tiles, Shard sh,
                      const int* live, int nlive, int num_sms, cudaStream_t stream) {
    using L = TmaSmem<KS, S>;
    static bool attr_set[64] = {};  // per device and instantiation: the attribute is set once, not per call
    int dev = 0;
    CSLAM_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        CSLAM_CUDA(cudaFuncSetAttribute(k_cov_update_tma<KS, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    const unsigned grid = (unsigned)std::min<long long>(num_sms, tiles);
    count_launch();
    if (g_tma_dense && g_tma_sub == 16) {
        constexpr int S16 = 2 * S > 8 ? 8 : 2 * S;  // same ring bytes with half-size stages (at most 8)
        static bool dense_attr[64] = {};
        if (dev < 0 || dev >= 64 || !dense_attr[dev]) {
            CSLAM_CUDA(cudaFuncSetAttribute(k_cov_update_tma_dense<KS, S16, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            TmdSmem<KS, S16, 16>::total));
            if (dev >= 0 && dev < 64) dense_attr[dev] = true;
        }
        k_cov_update_tma_dense<KS, S16, 16><<<grid, TMD_THREADS, TmdSmem<KS, S16, 16>::total, stream>>>(
            tm, tmd, A, lda, r, nt, tiles, sh, live, nlive, diag_eps, g_tma_dbg, g_tma_direct ? ds.P : nullptr, ds.ld, ds.rows);
    } else if (g_tma_dense) {
        static bool dense_attr[64] = {};
        if (dev < 0 || dev >= 64 || !dense_attr[dev]) {
            CSLAM_CUDA(cudaFuncSetAttribute(k_cov_update_tma_dense<KS, S, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            TmdSmem<KS, S, 32>::total));
            if (dev >= 0 && dev < 64) dense_attr[dev] = true;
        }
        k_cov_update_tma_dense<KS, S, 32><<<grid, TMD_THREADS, TmdSmem<KS, S, 32>::total, stream>>>(
            tm, tmd, A, lda, r, nt, tiles, sh, live, nlive, diag_eps, g_tma_dbg, g_tma_direct ? ds.P : nullptr, ds.ld, ds.rows);
    } else {
        k_cov_update_tma<KS, S><<<grid, TM_THREADS, L::total, stream>>>(tm, tmd, A, lda, r, nt, tiles, sh, live, nlive,
                                                                        diag_eps, g_tma_dbg, g_tma_boxr);
    }
    CSLAM_CUDA(cudaGetLastError());
    return CSLAM_OK;
}

// One pass P -= sum_{k < r} A[k] A[k]^T over this rank's tiles of the upper triangle (r <= 16).
// src_map / dst_map: tensor maps made by make_cov_tensor_map (the same one for an in-place pass).
int launch_cov_update_tma(const void* src_map, const void* dst_map, int n, const double* A, size_t lda, int r,
                          double diag_eps, Shard sh, const int* live, int nlive, int num_sms, int stages,
                          cudaStream_t stream, double* dst_ptr, size_t dst_ld, int dst_rows) {
    const DirectStore ds{dst_ptr, dst_ld, dst_rows};
    CSLAM_REQUIRE(r >= 1 && r <= 16, CSLAM_ERR_BAD_ARG, "TMA covariance pass: rank out of range (1..16)");
    CUtensorMap tm, tmd;
    memcpy(&tm, src_map, sizeof(tm));
    memcpy(&tmd, dst_map, sizeof(tmd));
    const int nt = (n + TM_T - 1) / TM_T;
    const long long tiles = shard_tile_count(nt, sh);
    if (tiles == 0) return CSLAM_OK;
    const int ks = (r + 3) / 4;
#define TM_CASE(KS_)                                                                                              \
    case KS_:                                                                                                     \
        if (stages == 4) return launch_one<KS_, 4>(tm, tmd, ds, diag_eps, A, lda, r, nt, tiles, sh, live, nlive, num_sms, stream);   \
        if (stages == 3) return launch_one<KS_, 3>(tm, tmd, ds, diag_eps, A, lda, r, nt, tiles, sh, live, nlive, num_sms, stream);   \
        if (stages == 2) return launch_one<KS_, 2>(tm, tmd, ds, diag_eps, A, lda, r, nt, tiles, sh, live, nlive, num_sms, stream);   \
        if (stages == 6 && KS_ <= 2)                                                                              \
            return launch_one<KS_, (KS_ <= 2 ? 6 : 5)>(tm, tmd, ds, diag_eps, A, lda, r, nt, tiles, sh, live, nlive, num_sms, stream); \
        return launch_one<KS_, 5>(tm, tmd, ds, diag_eps, A, lda, r, nt, tiles, sh, live, nlive, num_sms, stream);
    switch (ks) {
        TM_CASE(1) TM_CASE(2) TM_CASE(3) TM_CASE(4)
    }
#undef TM_CASE
    return CSLAM_ERR_BAD_ARG;
}

// Joint update P -= sum_{k < r} A[k] A[k]^T, r <= 64, in place on the array behind src_map (dense-row tensor map).
int launch_cov_update_tma_joint(const void* src_map, int n, const double* A, size_t lda, int r, Shard sh, int num_sms,
                                double* P, size_t ld, int rows_cap, cudaStream_t stream) {
    CSLAM_REQUIRE(r >= 1 && r <= 4 * TMJ_KS, CSLAM_ERR_BAD_ARG, "joint TMA pass: rank out of range (1..64)");
    CSLAM_REQUIRE(g_tma_dense && g_tma_sub == 32, CSLAM_ERR_UNSUPPORTED, "joint TMA pass needs the dense 32-row tensor map");
    CUtensorMap tm;
    memcpy(&tm, src_map, sizeof(tm));
    const int nt = (n + TM_T - 1) / TM_T;
    const long long tiles = shard_tile_count(nt, sh);
    if (tiles == 0) return CSLAM_OK;
    const unsigned grid = (unsigned)std::min<long long>(num_sms, tiles);
    static const int env_chunk = getenv("CSLAM_TMA_JOINT_CHUNK") ? atoi(getenv("CSLAM_TMA_JOINT_CHUNK")) : 0;
    const int chunk = env_chunk > 0 ? env_chunk : (int)std::max<long long>(1, std::min<long long>(8, tiles / ((long long)grid * 12)));
    using L = TmjSmem<2, 32>;
    static bool attr[64] = {};
    int dev = 0;
    CSLAM_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr[dev]) {
        CSLAM_CUDA(cudaFuncSetAttribute(k_cov_update_tma_joint<2, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total));
        if (dev >= 0 && dev < 64) attr[dev] = true;
    }
    count_launch();
    k_cov_update_tma_joint<2, 32><<<grid, TMJ_THREADS, L::total, stream>>>(tm, A, lda, r, nt, tiles, chunk, sh, P, ld, rows_cap);
    CSLAM_CUDA(cudaGetLastError());
    return CSLAM_OK;
}

}  // namespace cslam

// FP64 tensor-core peak of this GPU, measured here and now (MEASURED_PEAKS.json carries HBM and bf16 only): 16
// independent register-resident DMMA chains per warp, 16 warps per CTA, 4 CTAs per SM; best of 5 timed launches.
namespace cslam {
__global__ void __launch_bounds__(512) k_dmma_peak(double* out, int iters) {
    double c[16][2];
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
#pragma unroll
    for (int i = 0; i < 16; i++) c[i][0] = c[i][1] = 0.0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) ptx::dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;  // never true: keeps the chains alive
}
}  // namespace cslam

extern "C" int cslam_dmma_peak(int device, double* tflops) {
    CSLAM_REQUIRE(tflops != nullptr, CSLAM_ERR_BAD_ARG, "tflops is null");
    *tflops = 0.0;
    CSLAM_CUDA(cudaSetDevice(device));
    int sms = 0;
    CSLAM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    double* d = nullptr;
    CSLAM_CUDA(cudaMalloc(&d, 8));
    cudaEvent_t e0, e1;
    CSLAM_CUDA(cudaEventCreate(&e0));
    CSLAM_CUDA(cudaEventCreate(&e1));
    const int warps = 16, iters = 4000;
    const double flop = 2.0 * 8 * 8 * 4 * 16.0 * iters * warps * sms * 4.0;
    cslam::k_dmma_peak<<<sms, warps * 32>>>(d, 16);
    double best = 0.0;
    cudaError_t e = cudaDeviceSynchronize();
    for (int rep = 0; rep < 5 && e == cudaSuccess; rep++) {
        cslam::count_launch();
        cudaEventRecord(e0);
        cslam::k_dmma_peak<<<sms * 4, warps * 32>>>(d, iters);
        cudaEventRecord(e1);
        e = cudaEventSynchronize(e1);
        float ms = 0.f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        if (e == cudaSuccess && ms > 0.f) best = std::max(best, flop / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    CSLAM_CUDA(e);
    *tflops = best;
    return CSLAM_OK;
}
