// gate_parts.cuh — the per-block candidates of the gating kernel (gate.cu) and their final merge.
// k_gate leaves one (nd, j, outer) candidate per block and observation.  Its own last block can merge them
// (cslam_ekf_gate: the host wants the indices), or — fused scan on a deferred-pass handle — the column snapshot
// kernel that runs next merges them itself: every block needs the winner of ONE observation, a warp reads the
// ~150 candidates in one round trip, and the gate kernel's ticket / fence / last-block stage leaves the per-scan
// critical path.  The merge is a lexicographic (nd, j) minimum (EKF.cpp:235-326: strict '<', first landmark
// wins — SURVEY Q4), so the order of merging does not change the result.
#pragma once
#include "common.cuh"

namespace cslam {

struct Cand {
    double nd;
    double out;
    int j;
};
__device__ __forceinline__ void cand_merge(Cand& a, double nd, int j, double out) {
    if (nd < a.nd || (nd == a.nd && j < a.j)) {
        a.nd = nd;
        a.j = j;
    }
    if (out < a.out) a.out = out;
}

struct GateParts {
    const double* nd = nullptr;   // [nblocks][CSLAM_MAX_OBS]
    const double* out = nullptr;
    const int* j = nullptr;
    int nblocks = 0;
    int m = 0;                    // observations of the scan
    int* jbest = nullptr;         // final results (written by block (0, 0) of the snapshot kernel)
    double* nbest = nullptr;
    double* outer = nullptr;
    unsigned long long* assoc_count = nullptr;
};

// One warp: the final candidate of observation `obs` (all lanes return it).
__device__ __forceinline__ Cand gate_final_merge(const double* part_nd, const double* part_out, const int* part_j,
                                                 int nblocks, int obs, int lane) {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    Cand c{inf, inf, 0x7fffffff};
    // loads first (L2, ld.cg: the candidates were written by other blocks — of this kernel or the previous one),
    // merges after: one round trip per batch of 8 x 32 blocks, not one per candidate
    for (int b0 = 0; b0 < nblocks; b0 += 8 * 32) {
        double nd[8], out[8];
        int j[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int b = b0 + q * 32 + lane;
            const bool in = b < nblocks;
            const size_t at = (size_t)(in ? b : 0) * CSLAM_MAX_OBS + obs;
            nd[q] = in ? __ldcg(part_nd + at) : inf;
            out[q] = in ? __ldcg(part_out + at) : inf;
            j[q] = in ? __ldcg(part_j + at) : 0x7fffffff;
        }
#pragma unroll
        for (int q = 0; q < 8; q++) cand_merge(c, nd[q], j[q], out[q]);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double ond = __shfl_xor_sync(0xffffffffu, c.nd, off);
        const double oout = __shfl_xor_sync(0xffffffffu, c.out, off);
        const int oj = __shfl_xor_sync(0xffffffffu, c.j, off);
        cand_merge(c, ond, oj, oout);
    }
    return c;
}

}  // namespace cslam
