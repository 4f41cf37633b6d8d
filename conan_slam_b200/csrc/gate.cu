// gate.cu — Mahalanobis gating / nearest-neighbour association kernel (sm_100a).
// Separate translation unit: built with -fmad=false so that every +,-,*,/ and sqrt is the
// same IEEE operation, in the same order, as in oracle/slam_oracle.hpp
// (ekf_compute_association_sparse + PartialPivLU) — association indices must match exactly.
#include "common.cuh"
#include "gate_parts.cuh"

namespace cslam {

// ---------------------------------------------------------------------------- gating ----
// EKF.cpp:235-326 dataAssociate with computeAssociation (EKF.cpp:131-144) evaluated through
// the sparse H.  One thread per landmark: the 2x2 innovation covariance S_j, its inverse and
// log-determinant do not depend on the observation and are formed once, then reused for all
// m observations of the scan.  Per observation, a warp-shuffle lexicographic (nd, j) argmin
// over {j : nis < gate1} reproduces the reference's strict-'<' first-wins scan (Q4), followed
// by a per-block stage through shared memory and a last-block final reduction.
// Compiled with -fmad=false and written in the oracle's operation order so that nis / nd
// agree with the CPU restatement to the last bits (only atan2/log may differ by an ulp).
struct GatePack {
    double z[2 * CSLAM_MAX_OBS];
    int m;
    double R[4];
    double gate1, gate2;
};

// R3 = rows 0..2 of P (P itself on one GPU); D = replicated cache of the diagonal blocks
// ([3][dcap], sharded handles) or nullptr (read them from P).
constexpr int kGateThreads = 128;  // 4 warps: twice the CTAs of a 256-thread block (20k landmarks -> 157 CTAs >= 148 SMs)
__global__ void __launch_bounds__(kGateThreads) k_gate(const double* __restrict__ X, const double* __restrict__ P,
                                              const double* __restrict__ R3, const double* __restrict__ D, int dcap,
                                              size_t ld, int nf, GatePack gp, double* __restrict__ part_nd,
                                              double* __restrict__ part_out, int* __restrict__ part_j,
                                              unsigned* __restrict__ ticket, int* __restrict__ jbest,
                                              double* __restrict__ nbest, double* __restrict__ outer,
                                              unsigned long long* __restrict__ assoc_count, int final_stage) {
    constexpr int NW = kGateThreads / 32;
    __shared__ double s_nd[NW][CSLAM_MAX_OBS];
    __shared__ double s_out[NW][CSLAM_MAX_OBS];
    __shared__ int s_j[NW][CSLAM_MAX_OBS];
    __shared__ bool is_last;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const int jl = blockIdx.x * blockDim.x + threadIdx.x + 1;  // 1-based landmark id
    const bool valid = jl <= nf;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    double zhat_r = 0, zhat_b = 0, i00 = 0, i01 = 0, i10 = 0, i11 = 0, logdet = 0;
    if (valid) {
        const int f = 3 + 2 * (jl - 1);
        const int cols[5] = {0, 1, 2, f, f + 1};
        const ObsLin o = observe_lin(X[0], X[1], X[2], X[f], X[f + 1]);
        zhat_r = o.zr;
        zhat_b = o.zb;
        double H[2][5];
        for (int a = 0; a < 2; a++) {
            H[a][0] = o.hu[a][0]; H[a][1] = o.hu[a][1]; H[a][2] = o.hu[a][2];
            H[a][3] = o.lu[a][0]; H[a][4] = o.lu[a][1];
        }
        double Pc[5][5];
        for (int a = 0; a < 3; a++)
            for (int b = a; b < 5; b++) Pc[a][b] = Pc[b][a] = R3[(size_t)a * ld + cols[b]];
        if (D) {
            Pc[3][3] = D[jl - 1];
            Pc[3][4] = Pc[4][3] = D[(size_t)dcap + jl - 1];
            Pc[4][4] = D[2 * (size_t)dcap + jl - 1];
        } else {
            Pc[3][3] = P[(size_t)f * ld + f];
            Pc[3][4] = Pc[4][3] = P[(size_t)f * ld + f + 1];
            Pc[4][4] = P[(size_t)(f + 1) * ld + f + 1];
        }
        double HP[2][5];
        for (int a = 0; a < 2; a++)
            for (int b = 0; b < 5; b++) {
                double s = 0.0;
                for (int k = 0; k < 5; k++) s += H[a][k] * Pc[k][b];
                HP[a][b] = s;
            }
        double S[2][2];
        for (int a = 0; a < 2; a++)
            for (int b = 0; b < 2; b++) {
                double s = 0.0;
                for (int k = 0; k < 5; k++) s += HP[a][k] * H[b][k];
                S[a][b] = s + gp.R[a + 2 * b];
            }
        // 2x2 partial-pivot LU inverse and determinant, step for step as oracle::PartialPivLU
        double l00 = S[0][0], l01 = S[0][1], l10 = S[1][0], l11 = S[1][1];
        int p0 = 0, p1 = 1;
        double sign = 1.0;
        if (fabs(l10) > fabs(l00)) {
            double t;
            t = l00; l00 = l10; l10 = t;
            t = l01; l01 = l11; l11 = t;
            p0 = 1; p1 = 0;
            sign = -1.0;
        }
        l10 = l10 / l00;
        l11 = l11 - l10 * l01;
        const double det = (sign * l00) * l11;
        {
            const double y0 = (p0 == 0) ? 1.0 : 0.0;
            const double y1 = ((p1 == 0) ? 1.0 : 0.0) - l10 * y0;
            i10 = y1 / l11;
            i00 = (y0 - l01 * i10) / l00;
        }
        {
            const double y0 = (p0 == 1) ? 1.0 : 0.0;
            const double y1 = ((p1 == 1) ? 1.0 : 0.0) - l10 * y0;
            i11 = y1 / l11;
            i01 = (y0 - l01 * i11) / l00;
        }
        logdet = log(det);
    }
    for (int i = 0; i < gp.m; i++) {
        Cand c{inf, inf, 0x7fffffff};
        if (valid) {
            const double v0 = gp.z[2 * i] - zhat_r;
            const double v1 = pi2pi(gp.z[2 * i + 1] - zhat_b);
            const double t0 = v0 * i00 + v1 * i10;
            const double t1 = v0 * i01 + v1 * i11;
            const double nis = t0 * v0 + t1 * v1;
            const double nd = nis + logdet;
            if (nis < gp.gate1 && nd < inf) {
                c.nd = nd;
                c.j = jl;
            }
            if (nis < inf) c.out = nis;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double ond = __shfl_xor_sync(0xffffffffu, c.nd, off);
            const double oout = __shfl_xor_sync(0xffffffffu, c.out, off);
            const int oj = __shfl_xor_sync(0xffffffffu, c.j, off);
            cand_merge(c, ond, oj, oout);
        }
        if (lane == 0) {
            s_nd[warp][i] = c.nd;
            s_out[warp][i] = c.out;
            s_j[warp][i] = c.j;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < gp.m; i += blockDim.x) {
        Cand c{s_nd[0][i], s_out[0][i], s_j[0][i]};
        for (int w = 1; w < NW; w++) cand_merge(c, s_nd[w][i], s_j[w][i], s_out[w][i]);
        part_nd[(size_t)blockIdx.x * CSLAM_MAX_OBS + i] = c.nd;
        part_out[(size_t)blockIdx.x * CSLAM_MAX_OBS + i] = c.out;
        part_j[(size_t)blockIdx.x * CSLAM_MAX_OBS + i] = c.j;
    }
    if (!final_stage) return;  // the next kernel of the stream merges the candidates (gate_parts.cuh)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // final stage: one warp per observation, lanes stride over the per-block candidates, then the same
    // lexicographic (nd, j) merge across the warp — the order of merges does not change the result
    for (int i = warp; i < gp.m; i += NW) {
        const Cand c = gate_final_merge(part_nd, part_out, part_j, (int)gridDim.x, i, lane);
        if (lane == 0) {
            jbest[i] = (c.j == 0x7fffffff) ? 0 : c.j;
            nbest[i] = c.nd;
            outer[i] = c.out;
            // running total of associated observations: lets an asynchronous driver (cslam_ekf_scan with
            // no index read-back) learn how many updates its scans applied with ONE read at the end
            if (assoc_count != nullptr && c.j != 0x7fffffff) atomicAdd(assoc_count, 1ULL);
        }
    }
    if (threadIdx.x == 0) *ticket = 0;
}


int launch_gate(const double* X, const double* P, const double* R3, const double* D, int dcap, size_t ld, int nf,
                const double* Z, int m, const double R[4], double gate1, double gate2, double* part_nd,
                double* part_out, int* part_j, unsigned* ticket, int* jbest, double* nbest, double* outer,
                unsigned long long* assoc_count, cudaStream_t stream, int final_stage, int* nblocks_out) {
    GatePack gp;
    memset(&gp, 0, sizeof(gp));
    memcpy(gp.z, Z, sizeof(double) * 2 * m);
    gp.m = m;
    memcpy(gp.R, R, sizeof(double) * 4);
    gp.gate1 = gate1;
    gp.gate2 = gate2;
    const int blocks = nf > 0 ? (nf + kGateThreads - 1) / kGateThreads : 1;
    count_launch();
    k_gate<<<blocks, kGateThreads, 0, stream>>>(X, P, R3, D, dcap, ld, nf, gp, part_nd, part_out, part_j, ticket, jbest, nbest,
                                                outer, assoc_count, final_stage);
    CSLAM_CUDA(cudaGetLastError());
    if (nblocks_out) *nblocks_out = blocks;
    return CSLAM_OK;
}

}  // namespace cslam
