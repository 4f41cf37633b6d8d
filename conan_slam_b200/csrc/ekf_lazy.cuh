// ekf_lazy.cuh — deferred covariance passes for large and sharded maps (included by ekf.cu, same
// -fmad=false translation unit).  See LazyState in ekf_handle.cuh for the invariant:
//     P_dev(rows >= 3) = P_true + sum of the pending rank-1 terms,   X, R3 (rows 0..2), D (2x2 blocks) current.
// Every Kalman update of the reference is  P <- P - W1 W1^T  (slam.h:260; the Joseph form of the heading
// update slam.h:718 reduces to the same shape, see k_heading_gain), so a sequence of updates — the k heading
// updates of k control steps (EKF.cpp:328-352) and the m sequential landmark updates of a scan
// (EKF.cpp:457-479) — is a sum of rank-1 terms that ONE pass over the covariance can apply.  What an update
// needs of P_true is tiny: rows 0..2 (kept current eagerly), the observed landmark's two columns (read from a
// column snapshot and corrected by the pending terms inside the gain kernel) and, for gating, the landmarks'
// 2x2 diagonal blocks (kept current eagerly).  Nothing else of P is ever read between passes.
#pragma once
#include <algorithm>
#include <functional>

#include "common.cuh"
#include "cov_update.cuh"
#include "ekf_handle.cuh"
#include "gate_parts.cuh"

namespace cslam {

// defined in cov_tma.cu
int make_cov_tensor_map(void* out_map64, double* P, size_t ld, size_t rows);
int launch_cov_update_tma(const void* src_map, const void* dst_map, int n, const double* A, size_t lda, int r,
                          double diag_eps, Shard sh, const int* live, int nlive, int num_sms, int stages,
                          cudaStream_t stream, double* dst_ptr = nullptr, size_t dst_ld = 0, int dst_rows = 0);

int launch_cov_update_tma_joint(const void* src_map, int n, const double* A, size_t lda, int r, Shard sh, int num_sms,
                                double* P, size_t ld, int rows_cap, cudaStream_t stream);

constexpr int kSeqGroupLazy = kSeqGroupLazyMax;  // observations per column snapshot (2 x 8 columns exchanged at once)

// Pending rank-1 terms as a gain kernel sees them: first the rows of the bank that the pass in flight is
// applying (the column snapshot was taken from the array that pass READS), then the rows of the current bank.
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void ktrace_in(unsigned long long* tr) {
    if (tr != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) tr[0] = globaltimer_ns();
}
__device__ __forceinline__ void ktrace_out(unsigned long long* tr) {
    if (tr != nullptr && threadIdx.x == 0) atomicMax(tr + 1, globaltimer_ns());
}

// c - a * b with one rounding.  The chain's small kernels run beside a covariance pass that keeps the FP64 pipe
// busy with DMMAs: every instruction saved there is latency saved, and nothing requires the two-rounding form here
// (the pass itself accumulates fused).  Used by EVERY kernel that brings an entry up to date with pending terms, so
// that the fused group kernel, the one-kernel-per-observation chain and the follow kernel stay bit-identical.
__device__ __forceinline__ double nfma(double a, double b, double c) { return __fma_rn(-a, b, c); }

struct PendView {
    const double* A0;
    int n0;
    unsigned long long eps0;
    const double* A1;
    int n1;
    unsigned long long eps1;
    int r3_from;  // rows 0..2 are current up to the start of the running group: they lack terms >= r3_from only
};
__device__ __forceinline__ const double* pend_row(const PendView& pv, size_t lda, int t) {
    return t < pv.n0 ? pv.A0 + (size_t)t * lda : pv.A1 + (size_t)(t - pv.n0) * lda;
}
__device__ __forceinline__ bool pend_eps(const PendView& pv, int t) {
    return t < pv.n0 ? ((pv.eps0 >> t) & 1ull) != 0 : ((pv.eps1 >> (t - pv.n0)) & 1ull) != 0;
}

// What a whole group of sequential updates needs of the pending terms BEFORE it reads a single row: the pending
// panel rows at the columns of the group's marginal M = {0, 1, 2, f_0, f_0 + 1, ...}.  Block (0, 0) of the
// snapshot kernel gathers them next to its own work, so the group-gain kernel starts with ONE fixed-address read
// instead of "association index -> column -> pending rows" (two dependent round trips to a loaded memory system).
// (struct GroupHeader: ekf_handle.cuh)
// Fused scan, first snapshot of a scan: the gate kernel left per-block candidates (gate_parts.cuh).  Warp 0 of every
// block merges those of the observation its column belongs to; block (0, 0) merges all of them, publishes the
// scan's results (d_jbest / nbest / outer, running association count) and returns with s_j[k] = index of
// observation k for the header.  Returns the 1-based landmark index of observation `obs` (0 = none).
__device__ __forceinline__ int snapshot_associate(const GateParts& gp, int obs, bool lead_block, int* s_j) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (lead_block) {
        for (int i = warp; i < gp.m; i += nw) {
            const Cand c = gate_final_merge(gp.nd, gp.out, gp.j, gp.nblocks, i, lane);
            if (lane == 0) {
                const int j = (c.j == 0x7fffffff) ? 0 : c.j;
                s_j[i] = j;
                gp.jbest[i] = j;
                gp.nbest[i] = c.nd;
                gp.outer[i] = c.out;
                if (gp.assoc_count != nullptr && j != 0) atomicAdd(gp.assoc_count, 1ULL);
            }
        }
    } else if (warp == 0) {
        const Cand c = gate_final_merge(gp.nd, gp.out, gp.j, gp.nblocks, obs, lane);
        if (lane == 0) s_j[obs] = (c.j == 0x7fffffff) ? 0 : c.j;
    }
    __syncthreads();
    return s_j[obs];
}

__device__ __forceinline__ void write_group_header(GroupHeader* __restrict__ hdr, const PendView& pv, size_t lda,
                                                   const ColList& cl, const int* idf_dev) {
    const int g = cl.n >> 1, d = 3 + 2 * g, nterm = pv.n0 + pv.n1;
    __shared__ int s_f[kSeqGroupLazyMax];  // columns first (one read of the association indices), pending rows after
    if (threadIdx.x < g) {
        int c = cl.c[2 * threadIdx.x];
        if (idf_dev != nullptr) {
            const int j = idf_dev[threadIdx.x];
            c = j > 0 ? 3 + 2 * (j - 1) : -1;
        }
        s_f[threadIdx.x] = c;
        hdr->f[threadIdx.x] = c;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < nterm * d; idx += blockDim.x) {
        const int t = idx / d, b = idx % d;
        const int c = b < 3 ? b : (s_f[(b - 3) >> 1] >= 0 ? s_f[(b - 3) >> 1] + ((b - 3) & 1) : -1);
        hdr->Ac[t][b] = c >= 0 ? __ldcg(pend_row(pv, lda, t) + c) : 0.0;
    }
}

// Column snapshot: colbuf[k][i] = P_dev(i, c_k) for the listed columns (symmetric read of the upper triangle;
// rows 0..2 come from the always-current R3).  Sharded: every rank contributes what it stores, zeros
// elsewhere, and an all-reduce completes the columns (x + 0 is exact).  idf_dev (nullable): fused scan —
// column k belongs to observation k/2 whose 1-based landmark index sits in device memory (0 = none).
__global__ void __launch_bounds__(256) k_col_pack_lazy(const double* __restrict__ P, const double* __restrict__ R3,
                                                       size_t ld, int n, ColList cl, double* __restrict__ colbuf,
                                                       size_t lda, Shard sh, const int* idf_dev,
                                                       GroupHeader* __restrict__ hdr, PendView pv, GateParts gp,
                                                       unsigned long long* tr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    ktrace_in(tr);
    __shared__ int s_j[CSLAM_MAX_OBS];
    const bool lead = blockIdx.x == 0 && blockIdx.y == 0;
    int jmine = -1;
    if (gp.nd != nullptr) {
        jmine = snapshot_associate(gp, k >> 1, lead, s_j);
        idf_dev = s_j;  // (shared memory: generic pointer)
    }
    if (hdr != nullptr && lead) write_group_header(hdr, pv, lda, cl, idf_dev);
    if (tr != nullptr) {  // diagnostics only: every block reports when it is done
        __syncthreads();
        ktrace_out(tr);
    }
    if (i >= n) return;
    int c = cl.c[k];
    if (idf_dev != nullptr) {
        const int j = jmine >= 0 ? jmine : idf_dev[k >> 1];
        c = j > 0 ? 3 + 2 * (j - 1) + (k & 1) : -1;
    }
    double v = 0.0;
    if (c < 0) {
    } else if (i < 3) {
        if (sh.rank == 0) v = R3[(size_t)i * ld + c];
    } else if (i <= c) {
        if (shard_owns(sh, i)) v = P[shard_lrow(sh, i) * ld + c];
    } else {
        if (shard_owns(sh, c)) v = P[shard_lrow(sh, c) * ld + i];
    }
    colbuf[(size_t)k * lda + i] = v;
}

// Sharded column snapshot over NVLink peer memory — the exchange fused into the snapshot kernel, no collective:
// entry (i, c_k) of the symmetric P is stored by exactly one rank (rows 0..2: rank 0's always-current panel); that
// rank writes it into the snapshot buffer of EVERY rank (peer pointers from CUDA IPC, plain st.global over
// NVLink / NVSwitch).  The last block of the grid then publishes `epoch` in sig[rank] of every peer
// (fence.sys before the flag: the data is performed system-wide first).
struct PeerTab {
    double* xbuf[8];
    unsigned long long* sig[8];
};
__global__ void __launch_bounds__(256) k_col_push(const double* __restrict__ P, const double* __restrict__ R3, size_t ld,
                                                  int n, ColList cl, size_t lda, size_t buf_off, Shard sh,
                                                  const int* __restrict__ idf_dev, PeerTab pt,
                                                  unsigned long long epoch, unsigned* __restrict__ ticket,
                                                  GroupHeader* __restrict__ hdr, PendView pv, GateParts gp,
                                                  unsigned long long* tr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    ktrace_in(tr);
    __shared__ int s_j[CSLAM_MAX_OBS];
    const bool lead = blockIdx.x == 0 && blockIdx.y == 0;
    int jmine = -1;
    if (gp.nd != nullptr) {
        jmine = snapshot_associate(gp, k >> 1, lead, s_j);
        idf_dev = s_j;
    }
    if (hdr != nullptr && lead) write_group_header(hdr, pv, lda, cl, idf_dev);
    if (i < n) {
        int c = cl.c[k];
        if (idf_dev != nullptr) {
            const int j = jmine >= 0 ? jmine : idf_dev[k >> 1];
            c = j > 0 ? 3 + 2 * (j - 1) + (k & 1) : -1;
        }
        bool mine = false;
        double v = 0.0;
        if (c < 0) {
            mine = sh.rank == 0;  // unused column: rank 0 clears it everywhere
        } else if (i < 3) {
            mine = sh.rank == 0;
            if (mine) v = R3[(size_t)i * ld + c];
        } else if (i <= c) {
            mine = shard_owns(sh, i);
            if (mine) v = P[shard_lrow(sh, i) * ld + c];
        } else {
            mine = shard_owns(sh, c);
            if (mine) v = P[shard_lrow(sh, c) * ld + i];
        }
        if (mine) {
            const size_t off = buf_off + (size_t)k * lda + i;
            for (int p = 0; p < sh.world; p++) pt.xbuf[p][off] = v;
        }
    }
    // last block: every block's stores are performed system-wide (fence.sys + ticket), then raise the flags
    // (bar.sync orders the block's stores before thread 0's fence; the fence is cumulative — the pattern of a
    // cooperative-groups grid barrier: one system-scope fence per block, not one per thread)
    __shared__ bool is_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        is_last = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1;
    }
    __syncthreads();
    ktrace_out(tr);
    if (!is_last) return;
    __threadfence_system();
    if (threadIdx.x < sh.world) {
        volatile unsigned long long* s = pt.sig[threadIdx.x] + sh.rank;
        *s = epoch;
    }
    if (threadIdx.x == 0) *ticket = 0;
}
// The same exchange without any fence or flag (what the fused group-gain kernel reads): every entry travels as a
// 16-byte cell {low word, epoch, high word, epoch}.  Each 8-byte half carries its own epoch tag, so a reader that
// sees both tags equal to the epoch it waits for has the value, however the store was split on its way (the
// "LL" protocol of NCCL, with a 64-bit payload).  The sender neither fences nor signals — under a covariance pass
// that saturates the memory system a system-scope fence costs tens of microseconds — and the receiver's row threads
// poll exactly the cells they need.  Cells are reused two snapshots later (epoch parity); a rank can only be one
// snapshot ahead of its peers because its next gains need their pushes.
struct PeerLL {
    uint4* x[8];
};
__global__ void __launch_bounds__(256) k_col_push_ll(const double* __restrict__ P, const double* __restrict__ R3,
                                                     size_t ld, int n, ColList cl, size_t lda, size_t cell_off, Shard sh,
                                                     const int* idf_dev, PeerLL pt, unsigned epoch,
                                                     GroupHeader* __restrict__ hdr, PendView pv, GateParts gp,
                                                     unsigned long long* tr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    ktrace_in(tr);
    __shared__ int s_j[CSLAM_MAX_OBS];
    const bool lead = blockIdx.x == 0 && blockIdx.y == 0;
    int jmine = -1;
    if (gp.nd != nullptr) {
        jmine = snapshot_associate(gp, k >> 1, lead, s_j);
        idf_dev = s_j;
    }
    if (hdr != nullptr && lead) write_group_header(hdr, pv, lda, cl, idf_dev);
    if (i < n) {
        int c = cl.c[k];
        if (idf_dev != nullptr) {
            const int j = jmine >= 0 ? jmine : idf_dev[k >> 1];
            c = j > 0 ? 3 + 2 * (j - 1) + (k & 1) : -1;
        }
        bool mine = false;
        double v = 0.0;
        if (c < 0) {  // observation without a landmark: nobody reads the column
        } else if (i < 3) {
            mine = sh.rank == 0;
            if (mine) v = R3[(size_t)i * ld + c];
        } else if (i <= c) {
            mine = shard_owns(sh, i);
            if (mine) v = P[shard_lrow(sh, i) * ld + c];
        } else {
            mine = shard_owns(sh, c);
            if (mine) v = P[shard_lrow(sh, c) * ld + i];
        }
        if (mine) {
            const uint4 cell = make_uint4((unsigned)__double2loint(v), epoch, (unsigned)__double2hiint(v), epoch);
            const size_t off = cell_off + (size_t)k * lda + i;
            for (int p = 0; p < sh.world; p++) pt.x[p][off] = cell;
        }
    }
    if (tr != nullptr) {
        __syncthreads();
        ktrace_out(tr);
    }
}
__device__ __forceinline__ uint4 ll_load(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
// Bounded wait (~4 s) for one cell; a peer that never delivers is counted in `status` instead of hanging the GPU.
__device__ __forceinline__ double ll_wait(const uint4* p, uint4 v, unsigned epoch, int* status) {
    if (v.y != epoch || v.w != epoch) {
        const long long t0 = clock64();
        do {
            v = ll_load(p);
            if (clock64() - t0 > 8000000000LL) {
                atomicAdd(status, 1 << 20);
                break;
            }
        } while (v.y != epoch || v.w != epoch);
    }
    return __hiloint2double((int)v.z, (int)v.x);
}

// One warp: lane q waits until rank q has published `epoch` here.  A bounded wait: if a peer never arrives (a
// crashed rank) the kernel gives up after ~4 s and counts an error instead of hanging the GPU.
__global__ void k_wait_peers(const unsigned long long* sig, int world, unsigned long long epoch, int* __restrict__ status) {
    const int q = threadIdx.x;
    if (q >= world) return;
    const volatile unsigned long long* s = sig + q;
    const long long t0 = clock64();
    while (*s < epoch) {
        if (clock64() - t0 > 8000000000LL) {
            atomicAdd(status, 1 << 20);
            break;
        }
    }
    __threadfence_system();
}

// slam.h:243-259 for one observation (sparse H), reading P through the column snapshot and the pending terms.
// Writes Xout = Xin + W v and the two panel rows W1 (Aout, Aout + lda) that join the pending terms.
__global__ void __launch_bounds__(256) k_gain_lazy(const double* Xin, double* Xout, const double* __restrict__ R3,
                                                   const double* __restrict__ colbuf, size_t ld, size_t lda, int n,
                                                   double zr, double zb, int idf, double r00, double r10, double r01,
                                                   double r11, unsigned flags, PendView pv, double* Aout,
                                                   int* __restrict__ status, const int* __restrict__ idf_dev) {
    __shared__ GainSmall g;
    __shared__ double sPc[5][5];
    __shared__ double sAc[2 * kLazyBankMax][5];  // a_t[c] for the five columns c in {0, 1, 2, f, f+1}
    if (idf_dev != nullptr) idf = *idf_dev;
    const int i0 = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    if (idf == 0) {  // no landmark passed the gate: X is carried over, zero panel rows are a no-op update
        for (int i = i0; i < n; i += stride) {
            Xout[i] = Xin[i];
            Aout[i] = 0.0;
            Aout[lda + i] = 0.0;
        }
        return;
    }
    const int f = 3 + 2 * (idf - 1);
    const int nterm = pv.n0 + pv.n1;
    for (int idx = threadIdx.x; idx < nterm * 5; idx += blockDim.x) {
        const int t = idx / 5, b = idx % 5;
        sAc[t][b] = pend_row(pv, lda, t)[b < 3 ? b : f + (b - 3)];
    }
    __syncthreads();
    if (threadIdx.x < 25) {  // the 5x5 block of P_true at {0, 1, 2, f, f+1}
        const int a = threadIdx.x / 5, b = threadIdx.x % 5;
        const int lo = min(a, b), hi = max(a, b);  // block indices; columns: lo/hi < 3 ? itself : f + (x - 3)
        double v;
        int t0;
        if (lo < 3) {
            v = R3[(size_t)lo * ld + (hi < 3 ? hi : f + (hi - 3))];
            t0 = pv.r3_from;
        } else {  // both in {f, f+1}: P(f + lo - 3, f + hi - 3) = snapshot column (hi - 3), row f + lo - 3
            v = colbuf[(size_t)(hi - 3) * lda + f + (lo - 3)];
            t0 = 0;
        }
        for (int t = t0; t < nterm; t++) {
            v = nfma(sAc[t][lo], sAc[t][hi], v);
            if (lo == hi && pend_eps(pv, t)) v += kFltMin;
        }
        sPc[a][b] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double Pc[5][5];
        for (int a = 0; a < 5; a++)
            for (int b = 0; b < 5; b++) Pc[a][b] = sPc[a][b];
        const double R[4] = {r00, r10, r01, r11};
        gain_prologue(Xin, Pc, f, zr, zb, R, flags, g);
        if (!g.ok && blockIdx.x == 0) atomicAdd(status, 1);
    }
    __syncthreads();
    for (int i = i0; i < n; i += stride) {
        double pc[5];  // P_true(i, c) for c in {0, 1, 2, f, f+1}
#pragma unroll
        for (int b = 0; b < 3; b++) pc[b] = i <= b ? R3[(size_t)i * ld + b] : R3[(size_t)b * ld + i];
        pc[3] = colbuf[i];
        pc[4] = colbuf[lda + i];
        // the snapshot columns lack every pending term (rows 0..2 of them came from R3: only the group's own)
        for (int t = (i < 3 ? pv.r3_from : 0); t < pv.r3_from; t++) {
            const double ai = pend_row(pv, lda, t)[i];
            pc[3] = nfma(ai, sAc[t][3], pc[3]);
            pc[4] = nfma(ai, sAc[t][4], pc[4]);
            if (pend_eps(pv, t)) {
                if (i == f) pc[3] += kFltMin;
                if (i == f + 1) pc[4] += kFltMin;
            }
        }
        for (int t = pv.r3_from; t < nterm; t++) {
            const double ai = pend_row(pv, lda, t)[i];
#pragma unroll
            for (int b = 0; b < 5; b++) pc[b] = nfma(ai, sAc[t][b], pc[b]);
            if (pend_eps(pv, t)) {
                if (i < 3) pc[i] += kFltMin;
                if (i == f) pc[3] += kFltMin;
                if (i == f + 1) pc[4] += kFltMin;
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

// Rows 0..2 (R3) and the landmarks' diagonal blocks (D) follow `cnt` new rank-1 terms (rows Ab, Ab + lda, ...)
// in order: same operations on every rank, so the replicas stay bit-identical.  epsm bit q: term q is a
// heading update (diagonal += FLT_MIN, slam.h:719).
__global__ void __launch_bounds__(256) k_follow_lazy(double* __restrict__ R3, double* __restrict__ D, int dcap,
                                                     size_t ld, size_t lda, int n, int nf,
                                                     const double* __restrict__ Ab, int cnt, unsigned long long epsm) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int nrow = j < 3 ? j + 1 : 3;  // rows i <= j of the upper triangle
    double o[3];
    for (int i = 0; i < nrow; i++) o[i] = R3[(size_t)i * ld + j];
    const bool lm = j >= 3 && ((j - 3) & 1) == 0 && (j - 3) / 2 < nf;  // first coordinate of a landmark
    const int l = (j - 3) / 2;
    double d00 = 0.0, d01 = 0.0, d11 = 0.0;
    if (lm) {
        d00 = D[l];
        d01 = D[(size_t)dcap + l];
        d11 = D[2 * (size_t)dcap + l];
    }
    for (int q = 0; q < cnt; q++) {
        const double* a = Ab + (size_t)q * lda;
        const double aj = a[j];
        const bool eps = (epsm >> q) & 1ULL;
        for (int i = 0; i < nrow; i++) {
            o[i] = nfma(aj, a[i], o[i]);
            if (eps && i == j) o[i] += kFltMin;
        }
        if (lm) {
            const double aj1 = a[j + 1];
            d00 = nfma(aj, aj, d00);
            d01 = nfma(aj, aj1, d01);
            d11 = nfma(aj1, aj1, d11);
            if (eps) {
                d00 += kFltMin;
                d11 += kFltMin;
            }
        }
    }
    for (int i = 0; i < nrow; i++) R3[(size_t)i * ld + j] = o[i];
    if (lm) {
        D[l] = d00;
        D[(size_t)dcap + l] = d01;
        D[2 * (size_t)dcap + l] = d11;
    }
}

// ------------------------------------------------------------------------------------
// One launch for a whole group of sequential landmark updates (EKF.cpp:457-479: singleUpdate re-linearises after
// every observation) + what k_follow_lazy does + the wait for the peers' column pushes.
//
// The g updates of a group are sequentially dependent only through a SMALL system: update k needs the 5x5 block
// of P_true at {0,1,2,f_k,f_k+1} and X there, i.e. the marginal over M = pose + the group's landmarks (3 + 2g <= 19
// entries).  A row i of P_true restricted to the columns of M evolves under  P(i,:) -= W1_k(i) W1_k(:)^T  with
// W1_k(i) a function of that row alone and of the small quantities (H_k, G_k, W1_k at the rows of M).  So every
// block first replays the group on the |M| marginal rows (warp 0: lane a owns marginal row a, lane 0 factorises),
// keeping H_k, G_k, V_k and W1_k(M) in shared memory, and then every thread takes its own row through the g
// updates without any further synchronisation.  Each step performs exactly the operations k_gain_lazy /
// k_follow_lazy perform, in the same order, so the results are bit-identical to the one-kernel-per-observation
// chain (tests/test_ekf_lazy.py) — the launch count per scan drops from g + 2 to 1.
// R3 / D are written to their twins (the marginal phase of other blocks still reads the old ones).
struct ObsGroup {
    double z[2 * kSeqGroupLazyMax];
    int idf[kSeqGroupLazyMax];
    int g;
};

template <int GM>
struct GroupSmem {
    int f[GM];                                // first state index of observation k's landmark, -1 = skipped
    double Ac[2 * kLazyBankMax][3 + 2 * GM];      // pending terms (before the group) at the columns of M
    double W[2 * GM][3 + 2 * GM];              // the group's own panel rows at the rows of M
    GainSmall G[GM];
    double Pc[5][5];
    double X5[5];
};

template <int GM>
struct GroupRow {
    double pose[3];     // P_true(i, 0..2)
    double col[2 * GM]; // P_true(i, f_k), P_true(i, f_k + 1)
    double x;
};

template <int GM, bool LL>
__device__ __forceinline__ void group_row_load(GroupRow<GM>& r, int i, int g, const GroupSmem<GM>& sm,
                                               const double* Xin, const double* __restrict__ R3,
                                               const double* snap, size_t ld, size_t lda, const PendView& pv,
                                               unsigned epoch, int* status) {
#pragma unroll
    for (int b = 0; b < 3; b++) r.pose[b] = i <= b ? R3[(size_t)i * ld + b] : R3[(size_t)b * ld + i];
    if constexpr (LL) {  // flagged cells pushed by the peers: all loads first, then wait for the stragglers
        const uint4* cells = reinterpret_cast<const uint4*>(snap);
        uint4 v[2 * GM];
#pragma unroll
        for (int c = 0; c < 2 * GM; c++)
            if (c < 2 * g && sm.f[c >> 1] >= 0) v[c] = ll_load(cells + (size_t)c * lda + i);
#pragma unroll
        for (int c = 0; c < 2 * GM; c++)
            r.col[c] = (c < 2 * g && sm.f[c >> 1] >= 0) ? ll_wait(cells + (size_t)c * lda + i, v[c], epoch, status) : 0.0;
    } else {
#pragma unroll
        for (int c = 0; c < 2 * GM; c++)
            r.col[c] = (c < 2 * g && sm.f[c >> 1] >= 0) ? __ldcg(snap + (size_t)c * lda + i) : 0.0;
    }
    r.x = Xin[i];
    if (i < 3) return;  // rows 0..2 of the snapshot came from the always-current R3
    // pending terms in batches of 8: the loads of a batch are issued together (one round trip to a memory system
    // that the covariance pass keeps saturated — one-by-one they cost a full loaded latency each), the operations
    // stay in term order
    const int nterm = pv.n0 + pv.n1;
    for (int t0 = 0; t0 < nterm; t0 += 8) {
        double a[8];
#pragma unroll
        for (int q = 0; q < 8; q++) a[q] = t0 + q < nterm ? __ldcg(pend_row(pv, lda, t0 + q) + i) : 0.0;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int t = t0 + q;
            if (t < nterm) {
                const double ai = a[q];
                const bool eps = pend_eps(pv, t);
#pragma unroll
                for (int c = 0; c < 2 * GM; c++) {
                    if (c < 2 * g) {
                        r.col[c] = nfma(ai, sm.Ac[t][3 + c], r.col[c]);
                        if (eps && i == sm.f[c >> 1] + (c & 1) && sm.f[c >> 1] >= 0) r.col[c] += kFltMin;
                    }
                }
            }
        }
    }
}
// slam.h:257-259 for row i and observation k
template <int GM>
__device__ __forceinline__ void group_row_gain(GroupRow<GM>& r, int k, const GroupSmem<GM>& sm, double& w1_0,
                                               double& w1_1) {
    if (sm.f[k] < 0) {
        w1_0 = 0.0;
        w1_1 = 0.0;
        return;
    }
    const GainSmall& g = sm.G[k];
    double c0 = 0.0, c1 = 0.0;
#pragma unroll
    for (int q = 0; q < GM; q++)
        if (q == k) {
            c0 = r.col[2 * q];
            c1 = r.col[2 * q + 1];
        }
    double pht[2];
#pragma unroll
    for (int a = 0; a < 2; a++)
        pht[a] = (((r.pose[0] * g.H[a][0] + r.pose[1] * g.H[a][1]) + r.pose[2] * g.H[a][2]) + c0 * g.H[a][3]) + c1 * g.H[a][4];
    w1_0 = pht[0] * g.G[0][0] + pht[1] * g.G[1][0];
    w1_1 = pht[0] * g.G[0][1] + pht[1] * g.G[1][1];
    const double w_0 = w1_0 * g.G[0][0] + w1_1 * g.G[0][1];
    const double w_1 = w1_0 * g.G[1][0] + w1_1 * g.G[1][1];
    r.x = r.x + (w_0 * g.V[0] + w_1 * g.V[1]);
}
// slam.h:260 restricted to row i and the columns of M: the two panel rows of observation k, in order
template <int GM>
__device__ __forceinline__ void group_row_sub(GroupRow<GM>& r, int k, int g, const GroupSmem<GM>& sm, double w1_0,
                                              double w1_1) {
    if (sm.f[k] < 0) return;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const double ai = h == 0 ? w1_0 : w1_1;
        const double* w = sm.W[2 * k + h];
#pragma unroll
        for (int b = 0; b < 3; b++) r.pose[b] = nfma(ai, w[b], r.pose[b]);
#pragma unroll
        for (int c = 0; c < 2 * GM; c++)
            if (c < 2 * g) r.col[c] = nfma(ai, w[3 + c], r.col[c]);
    }
}

// 9 row warps + 1: at n = 40 003 the grid is 139 blocks <= 148 SMs (one wave), and 320 threads x 88 registers fit on
// an SM NEXT TO the persistent covariance pass (352 threads x 104 registers for a 16-row bank) — a block that had
// to wait for the pass to leave would stall the chain for the rest of the pass.
constexpr int kGroupRowThreads = 352;                    // warps 1..11: one state row each
constexpr int kGroupThreads = 32 + kGroupRowThreads;     // warp 0 replays the group on the marginal
template <int GM, bool LL>
__global__ void __maxnreg__(64) k_gain_group_lazy(
    const double* Xin, double* Xout, const double* __restrict__ R3, double* __restrict__ R3out,
    const double* __restrict__ D, double* __restrict__ Dout, int dcap, int nf, const double* snap, size_t ld, size_t lda,
    int n, ObsGroup og, double r00, double r10, double r01, double r11, unsigned flags, PendView pv, double* Aout,
    int* __restrict__ status, const GroupHeader* __restrict__ hdr, unsigned epoch, unsigned long long* tr) {
    __shared__ GroupSmem<GM> sm;
    ktrace_in(tr != nullptr ? tr + 2 : nullptr);
    const int g = og.g;
    const int tid = threadIdx.x;
    const int d = 3 + 2 * g;
    const int nterm = pv.n0 + pv.n1;
    // ---- phase 0: the header the snapshot kernel left (columns of M, pending terms there)
    if (tid < g) sm.f[tid] = hdr->f[tid];
    for (int idx = tid; idx < nterm * d; idx += blockDim.x) sm.Ac[idx / d][idx % d] = hdr->Ac[idx / d][idx % d];
    __syncthreads();
    if (tr != nullptr && blockIdx.x == 0 && tid == 0) tr[4] = globaltimer_ns();  // header + peers' flags in
    // rows are shifted by one so that the two coordinates of a landmark (rows 3 + 2l, 4 + 2l) sit in one warp on
    // lanes (even, odd); the row threads issue their loads BEFORE they wait for the marginal replay
    const int i = blockIdx.x * kGroupRowThreads + (tid - 32) - 1;
    const bool valid = tid >= 32 && i >= 0 && i < n;
    GroupRow<GM> r;
    if (tid < 32) {
        // ---- phase 1: the group replayed on the marginal rows (lane a owns row a of M, lane 0 factorises)
        const int a = tid;
        const bool act = a < d && (a < 3 || sm.f[(a - 3) >> 1] >= 0);
        const int ia = a < 3 ? a : (act ? sm.f[(a - 3) >> 1] + ((a - 3) & 1) : 0);
        if (act) group_row_load<GM, LL>(r, ia, g, sm, Xin, R3, snap, ld, lda, pv, epoch, status);
        for (int k = 0; k < g; k++) {
            if (sm.f[k] >= 0) {  // warp-uniform
                const int slot = a < 3 ? a : (a == 3 + 2 * k ? 3 : (a == 4 + 2 * k ? 4 : -1));
                if (slot >= 0 && act) {
                    double c0 = 0.0, c1 = 0.0;
#pragma unroll
                    for (int q = 0; q < GM; q++)
                        if (q == k) {
                            c0 = r.col[2 * q];
                            c1 = r.col[2 * q + 1];
                        }
                    sm.Pc[slot][0] = r.pose[0];
                    sm.Pc[slot][1] = r.pose[1];
                    sm.Pc[slot][2] = r.pose[2];
                    sm.Pc[slot][3] = c0;
                    sm.Pc[slot][4] = c1;
                    sm.X5[slot] = r.x;
                }
                __syncwarp();
                if (a == 0) {
                    double Pc[5][5], X5[5];
                    for (int p = 0; p < 5; p++) {
                        X5[p] = sm.X5[p];
                        for (int q = 0; q < 5; q++) Pc[p][q] = sm.Pc[p][q];
                    }
                    const double R[4] = {r00, r10, r01, r11};
                    gain_prologue(X5, Pc, 3, og.z[2 * k], og.z[2 * k + 1], R, flags, sm.G[k]);
                    if (!sm.G[k].ok && blockIdx.x == 0) atomicAdd(status, 1);
                }
                __syncwarp();
            }
            double w1_0 = 0.0, w1_1 = 0.0;
            if (act) group_row_gain<GM>(r, k, sm, w1_0, w1_1);
            if (a < d) {
                sm.W[2 * k][a] = w1_0;
                sm.W[2 * k + 1][a] = w1_1;
            }
            __syncwarp();
            // observation k is settled (H, G, V, W1 at the rows of M): release the row warps for it and go on with
            // k + 1 while they work — named barrier 1 + k, this warp arrives, the row warps wait
            __threadfence_block();
            asm volatile("barrier.arrive %0, %1;" ::"r"(1 + k), "r"(kGroupThreads) : "memory");
            if (act) group_row_sub<GM>(r, k, g, sm, w1_0, w1_1);
            __syncwarp();
        }
        if (tr != nullptr && blockIdx.x == 0 && tid == 0) tr[5] = globaltimer_ns();  // marginal replay done
        return;
    } else if (valid) {
        group_row_load<GM, LL>(r, i, g, sm, Xin, R3, snap, ld, lda, pv, epoch, status);
    }
    const bool lm = valid && i >= 3 && ((i - 3) & 1) == 0 && (i - 3) / 2 < nf;
    const int l = (i - 3) / 2;
    double d00 = 0.0, d01 = 0.0, d11 = 0.0;
    if (lm) {
        d00 = D[l];
        d01 = D[(size_t)dcap + l];
        d11 = D[2 * (size_t)dcap + l];
    }
    if (tr != nullptr && blockIdx.x == 0 && tid == 32) tr[6] = globaltimer_ns();  // row loads landed
    // ---- phase 2: every row thread takes its own row through the group, observation k as soon as warp 0 settled it
    for (int k = 0; k < g; k++) {
        asm volatile("barrier.sync %0, %1;" ::"r"(1 + k), "r"(kGroupThreads) : "memory");
        double w1_0 = 0.0, w1_1 = 0.0;
        if (valid) {
            group_row_gain<GM>(r, k, sm, w1_0, w1_1);
            group_row_sub<GM>(r, k, g, sm, w1_0, w1_1);
            Aout[(size_t)(2 * k) * lda + i] = w1_0;
            Aout[(size_t)(2 * k + 1) * lda + i] = w1_1;
        }
        const double n0 = __shfl_down_sync(0xffffffffu, w1_0, 1);
        const double n1 = __shfl_down_sync(0xffffffffu, w1_1, 1);
        if (lm) {
            d00 = nfma(w1_0, w1_0, d00);
            d01 = nfma(w1_0, n0, d01);
            d11 = nfma(n0, n0, d11);
            d00 = nfma(w1_1, w1_1, d00);
            d01 = nfma(w1_1, n1, d01);
            d11 = nfma(n1, n1, d11);
        }
    }
    if (valid) {
        Xout[i] = r.x;
        const int nrow = i < 3 ? i + 1 : 3;
        for (int b = 0; b < nrow; b++) R3out[(size_t)b * ld + i] = r.pose[b];
    }
    if (lm) {
        Dout[l] = d00;
        Dout[(size_t)dcap + l] = d01;
        Dout[2 * (size_t)dcap + l] = d11;
    }
    if (tr != nullptr && tid == 32) atomicMax(tr + 3, globaltimer_ns());
}

// ------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------
static inline double* lazy_bank(cslam_ekf* h, int bank) { return h->A + (size_t)bank * kLazyBankMax * h->lda; }
static inline double* lazy_P(cslam_ekf* h) { return h->lz.on ? h->lz.Pbuf[h->lz.stable] : h->P; }

// Launch one pass for the rows pending in the current bank (if any).  Chain stream: waits for the pass BEFORE
// this one (it produced the array that becomes `stable`, and it read the bank that becomes current).
static int lazy_flush(cslam_ekf* h) {
    LazyState& L = h->lz;
    if (!L.on || L.np == 0) return CSLAM_OK;
    // everything the chain has read from the array this pass overwrites must be done
    CSLAM_CUDA(cudaEventRecord(L.ev_chain, h->stream));
    CSLAM_CUDA(cudaStreamWaitEvent(L.pass_stream, L.ev_chain, 0));
    // One or two pending rows (a single heading or landmark update, e.g. a driver that flushes after every scan):
    // the FMA streaming kernel is the faster one there (1.98 ms = 99 % of HBM at N = 20k; the tensor-core pass pays
    // off from rank 4 on) — it works in place, so the chain must wait for it before it reads the array again.
    const bool small = L.np <= 2;
    const int src = L.newest, dst = (L.pingpong && !small) ? (L.newest ^ 1) : L.newest;
    const double eps = (double)__builtin_popcountll(L.eps_mask) * kFltMin;
    {
        ProfScope prof(h, L.pass_stream);
        if (small) {
            const int nt = (h->n + 127) / 128;
            const long long tiles = shard_tile_count(nt, h->sh);
            if (tiles > 0) {
                count_launch();
                if (L.np == 1)
                    k_cov_update<1, 128, 4, 4, 1><<<(unsigned)tiles, 256, 0, L.pass_stream>>>(
                        L.Pbuf[src], h->ld, h->n, lazy_bank(h, L.bank), h->lda, nt, eps, h->sh, nullptr);
                else
                    k_cov_update<2, 128, 4, 4, 1><<<(unsigned)tiles, 256, 0, L.pass_stream>>>(
                        L.Pbuf[src], h->ld, h->n, lazy_bank(h, L.bank), h->lda, nt, eps, h->sh, nullptr);
                CSLAM_CUDA(cudaGetLastError());
            }
        } else if (L.np <= kLazyBankTma) {
            if (int rc = launch_cov_update_tma(L.map[src], L.map[dst], h->n, lazy_bank(h, L.bank), h->lda, L.np, eps, h->sh,
                                               nullptr, 0, L.num_sms, L.stages, L.pass_stream, L.Pbuf[dst], h->ld,
                                               h->local_rows_cap))
                return rc;
        } else {
            // more than 16 rows: the FP64 tensor-core kernel of the joint update, reading `src` and writing `dst`
            if (int rc = launch_cov_update_dmma(L.Pbuf[src], h->ld, h->n, lazy_bank(h, L.bank), h->lda, L.np, h->sh,
                                                L.pass_panels, h->n_cap, 0, L.pass_stream, L.Pbuf[dst], eps))
                return rc;
        }
    }
    if (L.pass_pending_wait) CSLAM_CUDA(cudaStreamWaitEvent(h->stream, L.ev_pass, 0));  // the previous pass
    CSLAM_CUDA(cudaEventRecord(L.ev_pass, L.pass_stream));                               // ... now this one
    L.pass_pending_wait = true;
    L.stable = src;
    L.newest = dst;
    L.stable_busy = (dst == src);  // the pass in flight writes the very array the chain would read
    L.infl_rows = L.np;
    L.infl_eps_mask = L.eps_mask;
    L.bank ^= 1;
    L.np = 0;
    L.eps_mask = 0;
    L.passes++;
    return CSLAM_OK;
}

// Before the chain reads the covariance array: in-place mode has to wait for the pass in flight (it is
// modifying the array); with a ping-pong pair `stable` is never written while it is the read side.
static int lazy_acquire_read(cslam_ekf* h) {
    LazyState& L = h->lz;
    if (!L.on || !L.stable_busy || !L.pass_pending_wait) return CSLAM_OK;
    CSLAM_CUDA(cudaStreamWaitEvent(h->stream, L.ev_pass, 0));
    L.pass_pending_wait = false;
    L.stable_busy = false;
    L.infl_rows = 0;
    L.infl_eps_mask = 0;
    return CSLAM_OK;
}

// Apply everything: afterwards (in chain-stream order) lazy_P(h) holds P_true for rows >= 3.
static int lazy_flush_all(cslam_ekf* h) {
    LazyState& L = h->lz;
    if (!L.on) return CSLAM_OK;
    if (int rc = lazy_flush(h)) return rc;
    if (L.pass_pending_wait) {
        CSLAM_CUDA(cudaStreamWaitEvent(h->stream, L.ev_pass, 0));
        L.pass_pending_wait = false;
    }
    L.stable = L.newest;
    L.stable_busy = false;
    L.infl_rows = 0;
    L.infl_eps_mask = 0;
    return CSLAM_OK;
}

static int lazy_follow(cslam_ekf* h, const double* rows, int cnt, unsigned long long epsm) {
    const int nf = (h->n - 3) / 2;
    count_launch();
    k_follow_lazy<<<(h->n + 255) / 256, 256, 0, h->stream>>>(h->R3, h->D, h->dcap, h->ld, h->lda, h->n, nf, rows, cnt,
                                                            epsm);
    CSLAM_CUDA(cudaGetLastError());
    return CSLAM_OK;
}

// EKF.cpp:328-352 in lazy mode: the gain needs only rows 0..2; the rank-1 term joins the pending bank.
static int lazy_heading(cslam_ekf* h, double phi) {
    LazyState& L = h->lz;
    if (L.np + 1 > L.bank_rows)
        if (int rc = lazy_flush(h)) return rc;
    const int n = h->n;
    const double sigma = 0.01F * kPi / 180.0F;  // EKF.cpp:337
    double* row = lazy_bank(h, L.bank) + (size_t)L.np * h->lda;
    count_launch();
    k_heading_gain<<<(n + 255) / 256, 256, 0, h->stream>>>(h->X[h->cur], h->X[h->cur ^ 1], h->R3, h->ld, n, phi,
                                                           sigma * sigma, row);
    CSLAM_CUDA(cudaGetLastError());
    h->cur ^= 1;
    if (int rc = lazy_follow(h, row, 1, 1ULL)) return rc;
    L.eps_mask |= 1ull << L.np;
    L.np += 1;
    if (L.np == L.bank_rows) return lazy_flush(h);
    return CSLAM_OK;
}

static int allreduce_sum(cslam_ekf* h, double* buf, size_t count);

// Column snapshot of `ncols` columns (host list or device indices) from the array the chain may read.
// *snap receives the buffer the gains read ([ncols][lda]).
// *wait_fused (nullable): the caller's next kernel (k_gain_group_lazy) can wait for the peers' data itself; set to
// true when *snap points to flagged cells (k_col_push_ll) that it has to poll.
static int lazy_snapshot(cslam_ekf* h, const ColList& cl, const int* idf_dev, const double** snap,
                         bool* wait_fused = nullptr, const PendView* pvp = nullptr, const GateParts* gpp = nullptr) {
    if (int rc = lazy_acquire_read(h)) return rc;
    LazyState& L = h->lz;
    if (wait_fused) *wait_fused = false;
    GroupHeader* hdr = pvp ? L.hdr : nullptr;  // the fused group-gain kernel's header, gathered by block (0, 0)
    PendView pv{};
    if (pvp) pv = *pvp;
    GateParts gp;
    if (gpp) gp = *gpp;
    unsigned long long* tr = (pvp && L.ktrace) ? L.ktrace + 8 * (size_t)(L.kslot % LazyState::kTraceCap) : nullptr;
    if (h->sh.world > 1 && L.peers_ready && cl.n <= 2 * kSeqGroupLazyMax) {
        // peer-memory exchange: push what this rank stores into every rank's buffer
        L.epoch++;
        if (wait_fused) {  // flagged cells, no fence / flag / wait kernel: the reader polls the cells it needs
            const size_t cell_off = (size_t)(L.epoch & 1) * 2 * kSeqGroupLazyMax * h->lda;
            PeerLL pl;
            for (int q = 0; q < 8; q++)
                pl.x[q] = L.peer_xbuf[q] ? reinterpret_cast<uint4*>(reinterpret_cast<char*>(L.peer_xbuf[q]) + L.xll_off) : nullptr;
            count_launch();
            k_col_push_ll<<<dim3((h->n + 255) / 256, cl.n), 256, 0, h->stream>>>(lazy_P(h), h->R3, h->ld, h->n, cl, h->lda,
                                                                                 cell_off, h->sh, idf_dev, pl,
                                                                                 (unsigned)L.epoch, hdr, pv, gp, tr);
            CSLAM_CUDA(cudaGetLastError());
            *wait_fused = true;
            *snap = reinterpret_cast<const double*>(reinterpret_cast<const uint4*>(reinterpret_cast<char*>(L.xbuf) + L.xll_off) +
                                                    cell_off);
            return CSLAM_OK;
        }
        const size_t buf_off = (size_t)(L.epoch & 1) * 2 * kSeqGroupLazyMax * h->lda;
        PeerTab pt;
        for (int q = 0; q < 8; q++) {
            pt.xbuf[q] = L.peer_xbuf[q];
            pt.sig[q] = L.peer_sig[q];
        }
        count_launch();
        k_col_push<<<dim3((h->n + 255) / 256, cl.n), 256, 0, h->stream>>>(lazy_P(h), h->R3, h->ld, h->n, cl, h->lda, buf_off,
                                                                          h->sh, idf_dev, pt, L.epoch, L.push_ticket, hdr, pv, gp, tr);
        count_launch();
        k_wait_peers<<<1, 32, 0, h->stream>>>(L.sig, h->sh.world, L.epoch, h->status);
        CSLAM_CUDA(cudaGetLastError());
        *snap = L.xbuf + buf_off;
        return CSLAM_OK;
    }
    *snap = h->colbuf;
    count_launch();
    k_col_pack_lazy<<<dim3((h->n + 255) / 256, cl.n), 256, 0, h->stream>>>(lazy_P(h), h->R3, h->ld, h->n, cl, h->colbuf,
                                                                           h->lda, h->sh, idf_dev, hdr, pv, gp, tr);
    CSLAM_CUDA(cudaGetLastError());
    if (h->sh.world > 1) return allreduce_sum(h, h->colbuf, (size_t)cl.n * h->lda);
    return CSLAM_OK;
}

// singleUpdate (EKF.cpp:457-479) in lazy mode: per group of observations one column snapshot, per
// observation one gain kernel (re-linearised at the X the previous observation produced), then R3 / D
// follow; the 2g panel rows join the pending bank and a pass is launched whenever a bank is full.
// gate_parts (fused scan): the first snapshot merges the gate kernel's candidates; after_assoc queues the
// optional index read-back right behind that kernel.
static int lazy_sequential(cslam_ekf* h, const double* Z, const int32_t* idf_host, const int* idf_dev, int m,
                           const double R[4], const GateParts* gate_parts = nullptr,
                           const std::function<int()>* after_assoc = nullptr) {
    LazyState& L = h->lz;
    const int n = h->n;
    int base = 0;
    while (base < m) {
        if (L.bank_rows - L.np < 2)
            if (int rc = lazy_flush(h)) return rc;
        const int g = std::min({m - base, (L.bank_rows - L.np) / 2, kSeqGroupLazy});
        ColList cl;
        cl.n = 2 * g;
        for (int k = 0; k < g; k++) {
            cl.c[2 * k] = idf_host ? 3 + 2 * (idf_host[base + k] - 1) : 0;
            cl.c[2 * k + 1] = cl.c[2 * k] + 1;
        }
        const double* snap = nullptr;
        bool wait_fused = false;
        if (int rc = lazy_acquire_read(h)) return rc;  // in-place mode: the pass in flight first (changes the pending view)
        const int np0 = L.np;
        double* bank = lazy_bank(h, L.bank);
        PendView pvg;  // pending terms as they stand before the group
        pvg.A0 = lazy_bank(h, L.bank ^ 1);
        pvg.n0 = L.infl_rows;
        pvg.eps0 = L.infl_eps_mask;
        pvg.A1 = bank;
        pvg.n1 = np0;
        pvg.eps1 = L.eps_mask;
        pvg.r3_from = L.infl_rows + np0;
        if (int rc = lazy_snapshot(h, cl, idf_dev ? idf_dev + base : nullptr, &snap, L.fused_gains ? &wait_fused : nullptr,
                                   L.fused_gains ? &pvg : nullptr, base == 0 ? gate_parts : nullptr))
            return rc;
        if (base == 0 && after_assoc)
            if (int rc = (*after_assoc)()) return rc;
        if (L.fused_gains) {  // the whole group, the R3 / D follow and the wait for the peers in one launch
            const PendView& pv = pvg;
            ObsGroup og;
            og.g = g;
            for (int k = 0; k < kSeqGroupLazyMax; k++) {
                og.z[2 * k] = k < g ? Z[2 * (base + k)] : 0.0;
                og.z[2 * k + 1] = k < g ? Z[2 * (base + k) + 1] : 0.0;
                og.idf[k] = (k < g && idf_host) ? idf_host[base + k] : 0;
            }
            const unsigned blocks = (unsigned)((n + 1 + kGroupRowThreads - 1) / kGroupRowThreads);
            const int nf = (n - 3) / 2;
            count_launch();
#define CSLAM_GROUP(GM, LL)                                                                                            \
    k_gain_group_lazy<GM, LL><<<blocks, kGroupThreads, 0, h->stream>>>(                                                 \
        h->X[h->cur], h->X[h->cur ^ 1], h->R3, L.R3alt, h->D, L.Dalt, h->dcap, nf, snap, h->ld, h->lda, n, og, R[0], R[1], \
        R[2], R[3], h->flags, pv, bank + (size_t)np0 * h->lda, h->status, L.hdr, (unsigned)L.epoch,                     \
        L.ktrace ? L.ktrace + 8 * (size_t)(L.kslot++ % LazyState::kTraceCap) : nullptr)
#define CSLAM_GROUP2(GM)               \
    do {                               \
        if (wait_fused)                \
            CSLAM_GROUP(GM, true);     \
        else                           \
            CSLAM_GROUP(GM, false);    \
    } while (0)
            if (g <= 1) CSLAM_GROUP2(1);
            else if (g <= 2) CSLAM_GROUP2(2);
            else if (g <= 4) CSLAM_GROUP2(4);
            else CSLAM_GROUP2(8);
#undef CSLAM_GROUP2
#undef CSLAM_GROUP
            CSLAM_CUDA(cudaGetLastError());
            h->cur ^= 1;
            std::swap(h->R3, L.R3alt);
            std::swap(h->D, L.Dalt);
            L.np = np0 + 2 * g;
            base += g;
            if (L.np == L.bank_rows)
                if (int rc = lazy_flush(h)) return rc;
            continue;
        }
        for (int k = 0; k < g; k++) {
            const int i = base + k;
            PendView pv;
            pv.A0 = lazy_bank(h, L.bank ^ 1);
            pv.n0 = L.infl_rows;
            pv.eps0 = L.infl_eps_mask;
            pv.A1 = bank;
            pv.n1 = np0 + 2 * k;
            pv.eps1 = L.eps_mask;
            pv.r3_from = L.infl_rows + np0;
            count_launch();
            k_gain_lazy<<<(n + 255) / 256, 256, 0, h->stream>>>(
                h->X[h->cur], h->X[h->cur ^ 1], h->R3, snap + (size_t)2 * k * h->lda, h->ld, h->lda, n, Z[2 * i],
                Z[2 * i + 1], idf_host ? idf_host[i] : 0, R[0], R[1], R[2], R[3], h->flags, pv,
                bank + (size_t)(np0 + 2 * k) * h->lda, h->status, idf_dev ? idf_dev + i : nullptr);
            CSLAM_CUDA(cudaGetLastError());
            h->cur ^= 1;
        }
        if (int rc = lazy_follow(h, bank + (size_t)np0 * h->lda, 2 * g, 0ULL)) return rc;
        L.np = np0 + 2 * g;
        base += g;
        if (L.np == L.bank_rows)
            if (int rc = lazy_flush(h)) return rc;
    }
    return CSLAM_OK;
}

}  // namespace cslam
