// pf.cu — FastSLAM-style particle filter hot path for sm_100a (slam/src/PF.cpp) and the
// cslam_pf_* C ABI.  One thread per particle for the per-particle algebra (3x3 / 2x2,
// registers only), struct-of-arrays over particles so that every load/store of a warp is
// one contiguous 256-byte segment; stratified resampling = canonical radix-32
// warp-shuffle prefix scan + per-slot binary search + gather-copy (the bandwidth-dominant
// kernel: 2 x 40 B x Nf per particle).
//
// Layout in HBM (Pp = particle count rounded up to 32; two buffers A/B, the gather-copy
// ping-pongs between them):
//   w   [Pp]            weights
//   xv  [3][Pp]         pose (x, y, phi)
//   pv  [9][Pp]         pose covariance, full 3x3 row-major (the reference inverts it as a
//                       general matrix, PF.cpp:523-524)
//   xf  [2*Nf_cap][Pp]  feature means, row 2f+a
//   pf  [3*Nf_cap][Pp]  feature covariances, packed symmetric (xx, xy, yy), row 3f+b
//
// Built with -fmad=false and written in the oracle's operation order (matmul = ascending-k
// accumulation from zero, PartialPivLU inverse, LLT) so weights and poses track
// oracle/slam_oracle.hpp to rounding of libm calls only.
#include <algorithm>
#include <cstdlib>
#include <new>
#include <vector>

#include "common.cuh"
#include "nccl_dl.cuh"

namespace cslam {

struct PfBuf {
    double* w = nullptr;
    double* xv = nullptr;
    double* pv = nullptr;
    double* xf = nullptr;
    double* pf = nullptr;
};

}  // namespace cslam

struct cslam_pf {
    int device = 0;
    unsigned flags = 0;
    int np = 0;       // particles
    size_t pp = 0;    // padded stride
    int nf_cap = 0;
    int nf = 0;
    int cur = 0;
    cslam::PfBuf buf[2];
    double* comb = nullptr;    // [np] k/2 + i*k by repeated addition (PF.cpp:581-587), host-built
    double* scan[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // level-l inclusive scans
    size_t scan_len[5] = {0, 0, 0, 0, 0};
    int levels = 0;
    double* wn = nullptr;      // [np] normalised weights / cumulative weights
    int* keep = nullptr;       // [np]
    double* d_in = nullptr;    // staging for xi / u (3*np doubles)
    double* d_small = nullptr; // [8] scalars: sums, neff
    int* d_ismall = nullptr;   // [4] first-hit index, argmin, status
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // particles split over GPUs (world == 1: single GPU).  np = LOCAL particle count.
    int rank = 0, world = 1;
    long long np_global = 0;
    ncclComm_t comm = nullptr;
    int gather_level = -1;                 // first scan level computed on all-gathered totals
    double* gscan[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // global scan levels >= gather_level
    size_t gscan_len[5] = {0, 0, 0, 0, 0};
    double* cum_global = nullptr;          // [np_global] all-gathered cumulative weights
    cslam::PfBuf peer[2][8];               // IPC-mapped buffers of every rank (self = own pointers)
    bool peers_ready = false;
    const double** d_peer_tab = nullptr;   // device copy of the peers' base pointers
    double* acc_dev = nullptr;             // accessor scratch (AoS <-> SoA staging), grown on demand, never per call
    size_t acc_cap = 0;
    // diagnostics: CUDA events around the gather-copy launches of each resample
    bool prof = false;
    std::vector<cudaEvent_t> prof_ev;
    int prof_used = 0;
};

namespace cslam {

// PF.cpp:279-317 gaussEvaluate (likelihood branch)
template <int D>
__device__ __forceinline__ double gauss_evaluate(const double (&V)[D], const double (&S)[D][D], unsigned flags,
                                                 int* bad) {
    double L[D][D], SC[D][D], inv[D][D];
    if (!chol_lower<D>(S, L)) *bad = 1;
    tr<D, D>(L, SC);
    if (flags & CSLAM_FLAG_Q9_METRIC_S) inv_lu<D>(L, inv); else inv_lu<D>(SC, inv);
    double sum = 0.0;
#pragma unroll
    for (int i = 0; i < D; i++) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < D; k++) s += inv[i][k] * V[k];
        sum += s * s;
    }
    const double E = -0.5 * sum;
    double prod = 1.0;
#pragma unroll
    for (int i = 0; i < D; i++) prod *= SC[i][i];
    const double C = pow(2.0 * kPi, (double)D / 2.0) * prod;
    return exp(E) / C;
}

// PF.cpp:70-135 computeJacobians for one feature (bearing wrapped here)
struct Jac {
    double zp[2];
    double Hv[2][3];
    double Hf[2][2];
    double Sf[2][2];
};
__device__ __forceinline__ void compute_jacobians(const double (&X)[3], double fx, double fy, double pxx,
                                                  double pxy, double pyy, const double (&R)[2][2], Jac& J) {
    const double dx = fx - X[0], dy = fy - X[1];
    const double d2 = dx * dx + dy * dy, d = sqrt(d2);
    J.zp[0] = d;
    J.zp[1] = pi2pi(atan2(dy, dx) - X[2]);
    J.Hv[0][0] = -dx / d; J.Hv[0][1] = -dy / d; J.Hv[0][2] = 0.0;
    J.Hv[1][0] = dy / d2; J.Hv[1][1] = -dx / d2; J.Hv[1][2] = -1.0;
    J.Hf[0][0] = dx / d;   J.Hf[0][1] = dy / d;
    J.Hf[1][0] = -dy / d2; J.Hf[1][1] = dx / d2;
    const double Pf[2][2] = {{pxx, pxy}, {pxy, pyy}};
    double HP[2][2], HfT[2][2], HPH[2][2];
    mm<2, 2, 2>(J.Hf, Pf, HP);
    tr<2, 2>(J.Hf, HfT);
    mm<2, 2, 2>(HP, HfT, HPH);
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 2; b++) J.Sf[a][b] = HPH[a][b] + R[a][b];
}

struct PfObs {  // kernel-parameter transport of one scan
    double z[2 * CSLAM_MAX_OBS];
    int idf[CSLAM_MAX_OBS];
    int m;
    double R[2][2];
};

// ------------------------------------------------------------------------------------
// Per-particle kernels
// ------------------------------------------------------------------------------------

// PF.cpp:419-471 predict: P <- Gv P Gv^T + Gu Q Gu^T (3x3), deterministic pose advance (Q15).
__device__ __forceinline__ void pf_predict_regs(double (&X)[3], double (&P)[3][3], double v, double swa, double q00,
                                                double q01, double q10, double q11, double wb, double dt) {
    const double phi = X[2];
    const double s = sin(swa + phi), c = cos(swa + phi);
    const double Gv[3][3] = {{1.0, 0.0, -v * dt * s}, {0.0, 1.0, v * dt * c}, {0.0, 0.0, 1.0}};
    const double Gu[3][2] = {{dt * c, -v * dt * s}, {dt * s, v * dt * c}, {dt * sin(swa) / wb, v * dt * cos(swa) / wb}};
    const double Q[2][2] = {{q00, q01}, {q10, q11}};
    double GP[3][3], GvT[3][3], A[3][3], GQ[3][2], GuT[2][3], B[3][3];
    mm<3, 3, 3>(Gv, P, GP);
    tr<3, 3>(Gv, GvT);
    mm<3, 3, 3>(GP, GvT, A);
    mm<3, 2, 2>(Gu, Q, GQ);
    tr<3, 2>(Gu, GuT);
    mm<3, 2, 3>(GQ, GuT, B);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) P[i][j] = A[i][j] + B[i][j];
    const double x0 = X[0] + v * dt * c, x1 = X[1] + v * dt * s;
    X[2] = pi2pi(X[2] + v * dt * sin(swa) / wb);
    X[0] = x0;
    X[1] = x1;
}

// PF.cpp:382-417 observeHeading -> slam.h:700-725 josephUpdate on the 3x3 pose block, H = e_2^T.
__device__ __forceinline__ void pf_heading_regs(double (&X)[3], double (&P)[3][3], double phi_meas, double Rh) {
    const double v = pi2pi(phi_meas - X[2]);
    const double H[1][3] = {{0.0, 0.0, 1.0}};
    double Ht[3][1], PHT[3][1], HPHT[1][1];
    tr<1, 3>(H, Ht);
    mm<3, 3, 1>(P, Ht, PHT);
    mm<1, 3, 1>(H, PHT, HPHT);
    const double S = HPHT[0][0] + Rh;
    const double SI = (1.0 / S + 1.0 / S) * 0.5;  // inverse then makeSymmetric (1x1)
    double W[3];
#pragma unroll
    for (int i = 0; i < 3; i++) W[i] = PHT[i][0] * SI;
#pragma unroll
    for (int i = 0; i < 3; i++) X[i] = X[i] + W[i] * v;
    double Cm[3][3], WH[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            WH[i][j] = 0.0 + W[i] * H[0][j];
            Cm[i][j] = ((i == j) ? 1.0 : 0.0) - WH[i][j];
        }
    double CP[3][3], CmT[3][3], CPC[3][3];
    mm<3, 3, 3>(Cm, P, CP);
    tr<3, 3>(Cm, CmT);
    mm<3, 3, 3>(CP, CmT, CPC);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const double wr = 0.0 + W[i] * Rh;
            double o = CPC[i][j] + (0.0 + wr * W[j]);
            if (i == j) o = o + kFltMin;
            P[i][j] = o;
        }
}

__device__ __forceinline__ void pf_load_pose(const double* __restrict__ xv, const double* __restrict__ pv, size_t pp,
                                             int p, double (&X)[3], double (&P)[3][3]) {
#pragma unroll
    for (int i = 0; i < 3; i++) X[i] = xv[i * pp + p];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) P[i][j] = pv[(3 * i + j) * pp + p];
}
__device__ __forceinline__ void pf_store_pose(double* __restrict__ xv, double* __restrict__ pv, size_t pp, int p,
                                              const double (&X)[3], const double (&P)[3][3]) {
#pragma unroll
    for (int i = 0; i < 3; i++) xv[i * pp + p] = X[i];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) pv[(3 * i + j) * pp + p] = P[i][j];
}

__global__ void __launch_bounds__(256) k_pf_predict(double* __restrict__ xv, double* __restrict__ pv, size_t pp,
                                                    int np, double v, double swa, double q00, double q01,
                                                    double q10, double q11, double wb, double dt) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    double X[3], P[3][3];
    pf_load_pose(xv, pv, pp, p, X, P);
    pf_predict_regs(X, P, v, swa, q00, q01, q10, q11, wb, dt);
    pf_store_pose(xv, pv, pp, p, X, P);
}

__global__ void __launch_bounds__(256) k_pf_heading(double* __restrict__ xv, double* __restrict__ pv, size_t pp,
                                                    int np, double phi_meas, double Rh) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    double X[3], P[3][3];
    pf_load_pose(xv, pv, pp, p, X, P);
    pf_heading_regs(X, P, phi_meas, Rh);
    pf_store_pose(xv, pv, pp, p, X, P);
}

// k consecutive control steps of every particle (test/main.cpp:279-286 for k iterations of the driver
// loop): the 12 pose values are loaded once, k x (predict, observeHeading) run in registers, and are
// stored once — one launch and 2 x 104 B per particle instead of 2k launches and 2k x 208 B.
constexpr int kPfMaxControlSteps = 16;
struct PfControlPack {
    double v[kPfMaxControlSteps], swa[kPfMaxControlSteps], phi[kPfMaxControlSteps];
    int k, use_heading;
    double q00, q01, q10, q11, wb, dt, rh;
};
__global__ void __launch_bounds__(256) k_pf_control_steps(double* __restrict__ xv, double* __restrict__ pv, size_t pp,
                                                          int np, PfControlPack cp) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    double X[3], P[3][3];
    pf_load_pose(xv, pv, pp, p, X, P);
    for (int st = 0; st < cp.k; st++) {
        pf_predict_regs(X, P, cp.v[st], cp.swa[st], cp.q00, cp.q01, cp.q10, cp.q11, cp.wb, cp.dt);
        if (cp.use_heading) pf_heading_regs(X, P, cp.phi[st], cp.rh);
    }
    pf_store_pose(xv, pv, pp, p, X, P);
}

// PF.cpp:502-544 sampleProposal.
__global__ void __launch_bounds__(128) k_pf_sample_proposal(double* __restrict__ w, double* __restrict__ xv,
                                                            double* __restrict__ pv, const double* __restrict__ xf,
                                                            const double* __restrict__ pf, size_t pp, int np,
                                                            PfObs ob, const double* __restrict__ xi,
                                                            unsigned flags, int* __restrict__ status) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    int bad = 0;
    double X[3], P[3][3], X0[3], P0[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++) X0[i] = X[i] = xv[i * pp + p];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) P0[i][j] = P[i][j] = pv[(3 * i + j) * pp + p];
    for (int k = 0; k < ob.m; k++) {
        const int f = ob.idf[k] - 1;
        Jac J;
        compute_jacobians(X, xf[(size_t)(2 * f) * pp + p], xf[(size_t)(2 * f + 1) * pp + p],
                          pf[(size_t)(3 * f) * pp + p], pf[(size_t)(3 * f + 1) * pp + p],
                          pf[(size_t)(3 * f + 2) * pp + p], ob.R, J);
        double Sfi[2][2];
        inv_lu<2>(J.Sf, Sfi);
        const double V[2][1] = {{ob.z[2 * k] - J.zp[0]}, {pi2pi(ob.z[2 * k + 1] - J.zp[1])}};
        double HvT[3][2], HS[3][2], HSH[3][3], Pinv[3][3], PT[3][3];
        tr<2, 3>(J.Hv, HvT);
        mm<3, 2, 2>(HvT, Sfi, HS);
        mm<3, 2, 3>(HS, J.Hv, HSH);
        inv_lu<3>(P, Pinv);
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) PT[i][j] = HSH[i][j] + Pinv[i][j];
        inv_lu<3>(PT, P);
        double PH[3][2], PHS[3][2], dX[3][1];
        mm<3, 3, 2>(P, HvT, PH);
        mm<3, 2, 2>(PH, Sfi, PHS);
        mm<3, 2, 1>(PHS, V, dX);
#pragma unroll
        for (int i = 0; i < 3; i++) X[i] = X[i] + dX[i][0];
    }
    // slam.h:753-764: XS = chol(P) * xi + X
    double Lp[3][3], XS[3];
    if (!chol_lower<3>(P, Lp)) bad = 1;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 3; k++) s += Lp[i][k] * xi[(size_t)3 * p + k];
        XS[i] = s + X[i];
    }
    // likelihood at the sampled pose (PF.cpp:343-359)
    double like = 1.0;
    for (int k = 0; k < ob.m; k++) {
        const int f = ob.idf[k] - 1;
        Jac J;
        compute_jacobians(XS, xf[(size_t)(2 * f) * pp + p], xf[(size_t)(2 * f + 1) * pp + p],
                          pf[(size_t)(3 * f) * pp + p], pf[(size_t)(3 * f + 1) * pp + p],
                          pf[(size_t)(3 * f + 2) * pp + p], ob.R, J);
        const double V[2] = {ob.z[2 * k] - J.zp[0], pi2pi(ob.z[2 * k + 1] - J.zp[1])};
        like = like * gauss_evaluate<2>(V, J.Sf, flags, &bad);
    }
    const double d0[3] = {X0[0] - XS[0], X0[1] - XS[1], pi2pi(X0[2] - XS[2])};
    const double d1[3] = {X[0] - XS[0], X[1] - XS[1], pi2pi(X[2] - XS[2])};
    const double prior = gauss_evaluate<3>(d0, P0, flags, &bad);
    const double prop = gauss_evaluate<3>(d1, P, flags, &bad);
    w[p] = w[p] * like * prior / prop;
#pragma unroll
    for (int i = 0; i < 3; i++) xv[i * pp + p] = XS[i];
#pragma unroll
    for (int i = 0; i < 9; i++) pv[i * pp + p] = 0.0;
    if (bad) atomicAdd(status, 1);
}

// PF.cpp:222-277 featureUpdate: slam.h:235-266 choleskyUpdate on each observed 2x2 feature.
__global__ void __launch_bounds__(128) k_pf_feature_update(const double* __restrict__ xv, double* __restrict__ xf,
                                                           double* __restrict__ pf, size_t pp, int np, PfObs ob,
                                                           unsigned flags, int* __restrict__ status) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    int bad = 0;
    const double X[3] = {xv[p], xv[pp + p], xv[2 * pp + p]};
    for (int k = 0; k < ob.m; k++) {
        const int f = ob.idf[k] - 1;
        double fx = xf[(size_t)(2 * f) * pp + p], fy = xf[(size_t)(2 * f + 1) * pp + p];
        const double pxx = pf[(size_t)(3 * f) * pp + p], pxy = pf[(size_t)(3 * f + 1) * pp + p],
                     pyy = pf[(size_t)(3 * f + 2) * pp + p];
        Jac J;
        compute_jacobians(X, fx, fy, pxx, pxy, pyy, ob.R, J);
        const double V[2] = {ob.z[2 * k] - J.zp[0], pi2pi(ob.z[2 * k + 1] - J.zp[1])};
        // choleskyUpdate with n = 2, r = 2, H = Hf
        const double Pf[2][2] = {{pxx, pxy}, {pxy, pyy}};
        double HfT[2][2], PHT[2][2], HPHT[2][2], S[2][2];
        tr<2, 2>(J.Hf, HfT);
        mm<2, 2, 2>(Pf, HfT, PHT);
        mm<2, 2, 2>(J.Hf, PHT, HPHT);
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int b = 0; b < 2; b++) S[a][b] = HPHT[a][b] + ob.R[a][b];
        double Ss[2][2];
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int b = 0; b < 2; b++) Ss[a][b] = (S[a][b] + S[b][a]) * 0.5;
        double L[2][2], Li[2][2];
        const bool okc = chol_lower<2>(Ss, L);
        inv_lu<2>(L, Li);
        bool fin = okc;
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int b = 0; b < 2; b++) fin = fin && isfinite(Li[a][b]);
        if (!fin) {
            bad = 1;
#pragma unroll
            for (int a = 0; a < 2; a++)
#pragma unroll
                for (int b = 0; b < 2; b++) Li[a][b] = 0.0;
        }
        double G[2][2], GT[2][2];
        if (flags & CSLAM_FLAG_Q1_METRIC_S) { tr<2, 2>(Li, G); } else {
#pragma unroll
            for (int a = 0; a < 2; a++)
#pragma unroll
                for (int b = 0; b < 2; b++) G[a][b] = Li[a][b];
        }
        tr<2, 2>(G, GT);
        double W1[2][2], W[2][2], W1T[2][2], WW[2][2];
        mm<2, 2, 2>(PHT, G, W1);
        mm<2, 2, 2>(W1, GT, W);
        fx = fx + ((0.0 + W[0][0] * V[0]) + W[0][1] * V[1]);
        fy = fy + ((0.0 + W[1][0] * V[0]) + W[1][1] * V[1]);
        tr<2, 2>(W1, W1T);
        mm<2, 2, 2>(W1, W1T, WW);
        xf[(size_t)(2 * f) * pp + p] = fx;
        xf[(size_t)(2 * f + 1) * pp + p] = fy;
        pf[(size_t)(3 * f) * pp + p] = pxx - WW[0][0];
        pf[(size_t)(3 * f + 1) * pp + p] = pxy - WW[0][1];
        pf[(size_t)(3 * f + 2) * pp + p] = pyy - WW[1][1];
    }
    if (bad) atomicAdd(status, 1);
}

// PF.cpp:9-60 addOneNewFeature (particle overload)
__global__ void __launch_bounds__(256) k_pf_add_features(const double* __restrict__ xv, double* __restrict__ xf,
                                                         double* __restrict__ pf, size_t pp, int np, int nf0,
                                                         PfObs ob) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    const double X[3] = {xv[p], xv[pp + p], xv[2 * pp + p]};
    for (int i = 0; i < ob.m; i++) {
        const double r = ob.z[2 * i], b = ob.z[2 * i + 1];
        const double s = sin(X[2] + b), c = cos(X[2] + b);
        const int f = nf0 + i;
        xf[(size_t)(2 * f) * pp + p] = X[0] + r * c;
        xf[(size_t)(2 * f + 1) * pp + p] = X[1] + r * s;
        const double Gz[2][2] = {{c, -r * s}, {s, r * c}};
        double GR[2][2], GzT[2][2], Pn[2][2];
        mm<2, 2, 2>(Gz, ob.R, GR);
        tr<2, 2>(Gz, GzT);
        mm<2, 2, 2>(GR, GzT, Pn);
        pf[(size_t)(3 * f) * pp + p] = Pn[0][0];
        pf[(size_t)(3 * f + 1) * pp + p] = Pn[0][1];
        pf[(size_t)(3 * f + 2) * pp + p] = Pn[1][1];
    }
}

// test/main.cpp:319-325: X <- chol(P) xi + X ; P <- 0
__global__ void __launch_bounds__(256) k_pf_sample_pose(double* __restrict__ xv, double* __restrict__ pv, size_t pp,
                                                        int np, const double* __restrict__ xi,
                                                        int* __restrict__ status) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    double X[3], P[3][3], L[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++) X[i] = xv[i * pp + p];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) P[i][j] = pv[(3 * i + j) * pp + p];
    chol_lower<3>(P, L);  // an all-zero P (fresh particles) legitimately yields the zero factor
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 3; k++) s += L[i][k] * xi[(size_t)3 * p + k];
        xv[i * pp + p] = s + X[i];
    }
#pragma unroll
    for (int i = 0; i < 9; i++) pv[i * pp + p] = 0.0;
}

// ------------------------------------------------------------------------------------
// Resampling (PF.cpp:473-500, 546-596)
//
// Canonical summation order ("radix-32 hierarchical Kogge-Stone"), shared with the oracle's
// INTENDED mode so that cumulative weights — hence resampled indices — are bit-identical on
// 1, 2, 4 or 8 GPUs:
//   level 0: every aligned group of 32 consecutive values gets an inclusive Kogge-Stone
//            scan (x_i += x_{i-d}, d = 1,2,4,8,16: exactly a __shfl_up warp scan);
//   level l: the group totals of level l-1 are scanned the same way in groups of 32;
//   prefix(i) = scan_0[i] + sum over levels l >= 1 of the exclusive group offset, added
//            from the TOP level down:  off = (((0 + o_top) + ...) + o_1);  cum = off + scan_0[i].
// For np <= 32 this is NOT the reference's left-to-right order; REF_LITERAL mode therefore
// runs the plain sequential sum (one thread; it is only meaningful for tiny particle sets,
// since the literal resampler collapses to one particle anyway).
// ------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_scan_incl(double x, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x = x + y;
    }
    return x;
}

// out[i] = inclusive scan within aligned 32-groups of f(in[i]); totals[g] = last of group.
// mode 0: f(x) = x ; mode 1: f(x) = x / div[0] ; mode 2: f(x) = (x / div[0])^2 — the divisor
// lives on the device so no host round trip separates the passes.
// (in may alias out: upper levels are scanned in place)
__global__ void __launch_bounds__(256) k_scan_level(const double* in, size_t len, double* out, double* totals,
                                                    int mode, const double* __restrict__ div) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    double x = 0.0;
    if (i < len) {
        x = in[i];
        if (mode >= 1) x = x / div[0];
        if (mode == 2) x = x * x;
    }
    x = warp_scan_incl(x, lane);
    if (i < len) out[i] = x;
    const size_t g = i >> 5;
    const size_t last = ((g << 5) + 31 < len) ? (g << 5) + 31 : len - 1;
    if (i == last) totals[g] = x;
}

// total (canonical order) of a scanned hierarchy = last element of the top level
__global__ void k_scan_total(const double* __restrict__ top, size_t top_len, double* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = top[top_len - 1];
}

// sequential left-to-right sum (REF_LITERAL): out[0] = sum f(in[i])
__global__ void k_seq_sum(const double* __restrict__ in, size_t len, int mode, const double* __restrict__ div,
                          double* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s = 0.0;
    for (size_t i = 0; i < len; i++) {
        double x = in[i];
        if (mode >= 1) x = x / div[0];
        if (mode == 2) x = x * x;
        s += x;
    }
    out[0] = s;
}
__global__ void k_seq_cumsum(const double* __restrict__ in, size_t len, const double* __restrict__ div,
                             double* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s = in[0] / div[0];
    out[0] = s;
    for (size_t i = 1; i < len; i++) {
        s = s + in[i] / div[0];
        out[i] = s;
    }
}

// w[i] /= div[0]  (PF.cpp:484-487)
__global__ void __launch_bounds__(256) k_divide(double* __restrict__ w, size_t len, const double* __restrict__ div) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < len) w[i] = w[i] / div[0];
}

// cum[i] = offsets (top level down) + scan0[i].  Levels below `glevel` are local arrays indexed by the
// local group index; levels >= glevel are the all-gathered (global) arrays indexed by the global one.
struct ScanPtrs {
    const double* s[5];
    int levels;
    int glevel;          // == levels when nothing is gathered (single GPU)
    long long base;      // global index of local element 0 (rank * np_local)
};
__global__ void __launch_bounds__(256) k_scan_combine(ScanPtrs sp, size_t len, double* __restrict__ cum) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    double off = 0.0;
    for (int l = sp.levels - 1; l >= 1; l--) {
        const size_t gi = (l >= sp.glevel) ? (size_t)((sp.base + (long long)i) >> (5 * l)) : (i >> (5 * l));
        if (gi & 31) off = off + sp.s[l][gi - 1];
    }
    cum[i] = off + sp.s[0][i];
}

// PF.cpp:566-574.  INTENDED (Q10): keep[c] = min{ i : select[c] < cum[i] } by binary search.
// `cum` holds ncum cumulative weights (all particles of all ranks); this rank resolves its nslots slots.
__global__ void __launch_bounds__(256) k_resample_search(const double* __restrict__ cum, int ncum,
                                                         const double* __restrict__ comb,
                                                         const double* __restrict__ u, int nslots,
                                                         int* __restrict__ keep) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nslots) return;
    const double k = 1.0 / (double)ncum;
    const double sel = comb[c] + u[c] * (k - k / 2.0);
    int lo = 0, hi = ncum;
    while (lo < hi) {
        const int mid = (lo + hi) / 2;
        if (sel < cum[mid]) hi = mid; else lo = mid + 1;
    }
    keep[c] = lo < ncum ? lo : ncum - 1;
}
// REF_LITERAL: the first i with select[i] < cum[i] takes every slot (all zeros if none).
__global__ void __launch_bounds__(256) k_resample_first_hit(const double* __restrict__ cum,
                                                            const double* __restrict__ comb,
                                                            const double* __restrict__ u, int np,
                                                            int* __restrict__ first) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const double k = 1.0 / (double)np;
    const double sel = comb[i] + u[i] * (k - k / 2.0);
    if (sel < cum[i]) atomicMin(first, i);
}
__global__ void __launch_bounds__(256) k_fill_keep(int* __restrict__ keep, int np, const int* __restrict__ first) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= np) return;
    const int f = *first;
    keep[c] = (f >= np) ? 0 : f;
}

// Gather-copy of the surviving particles: dst[row][c] = src[row][keep[c]] for every SoA row.
// One thread owns TWO adjacent slots (one 16-byte store per row) and walks grid.y-strided rows;
// keep[] is ascending for a stratified comb, so the two 8-byte gathers of a warp fall in a
// handful of sectors.  This kernel moves 2 x 40 B x Nf per particle — the PF's HBM roofline.
__global__ void __launch_bounds__(256) k_gather_rows(const double* __restrict__ src, double* __restrict__ dst,
                                                     size_t pp, int np, int rows, const int* __restrict__ keep) {
    const int c = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (c >= np) return;
    const int k0 = keep[c];
    const int k1 = (c + 1 < np) ? keep[c + 1] : k0;
    for (int r = blockIdx.y; r < rows; r += gridDim.y) {
        const double* s = src + (size_t)r * pp;
        double2 v;
        v.x = s[k0];
        v.y = s[k1];
        st128(dst + (size_t)r * pp + c, v);
    }
}
// Multi-GPU variant: keep[] holds GLOBAL source indices; the survivor's rows are read straight from
// the owning rank's buffer over NVLink (IPC-mapped peer pointers) — the all-to-all of the
// resampling step is fused into the gather kernel, no staging copy.
// (the per-rank base pointers live in a small device table: indexing a by-value kernel parameter
//  dynamically would spill it to local memory in every thread)
__global__ void __launch_bounds__(256) k_gather_rows_peer(const double* const* __restrict__ base,
                                                          double* __restrict__ dst, size_t pp, int np, int rows,
                                                          const int* __restrict__ keep) {
    const int c = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (c >= np) return;
    const int g0 = keep[c];
    const int g1 = (c + 1 < np) ? keep[c + 1] : g0;
    const double* __restrict__ s0 = base[g0 / np] + (g0 % np);
    const double* __restrict__ s1 = base[g1 / np] + (g1 % np);
    for (int r = blockIdx.y; r < rows; r += gridDim.y) {
        double2 v;
        v.x = s0[(size_t)r * pp];
        v.y = s1[(size_t)r * pp];
        st128(dst + (size_t)r * pp + c, v);
    }
}
__global__ void __launch_bounds__(256) k_fill(double* __restrict__ a, size_t len, double v) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < len) a[i] = v;
}

// slam.h:493-511: first minimum-weight particle (Q13).  Stage 1: one (w, index) candidate per
// warp; stage 2: a single block reduces the candidates lexicographically.
__device__ __forceinline__ void argmin_merge(double& x, int& idx, double ox, int oi) {
    if (ox < x || (ox == x && oi < idx)) { x = ox; idx = oi; }
}
__global__ void __launch_bounds__(256) k_argmin_w(const double* __restrict__ w, int np, double* __restrict__ cw,
                                                  int* __restrict__ ci) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double x = (i < np) ? w[i] : __longlong_as_double(0x7ff0000000000000LL);
    int idx = (i < np) ? i : 0x7fffffff;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double ox = __shfl_xor_sync(0xffffffffu, x, off);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, off);
        argmin_merge(x, idx, ox, oi);
    }
    if ((threadIdx.x & 31) == 0 && i < np) {
        cw[i >> 5] = x;
        ci[i >> 5] = idx;
    }
}
__global__ void __launch_bounds__(1024) k_argmin_final(const double* __restrict__ cw, const int* __restrict__ ci,
                                                       int nw, const double* __restrict__ xv, size_t pp,
                                                       double* __restrict__ out /* [4]: x y phi index */) {
    __shared__ double sw[32];
    __shared__ int si[32];
    double x = __longlong_as_double(0x7ff0000000000000LL);
    int idx = 0x7fffffff;
    for (int i = threadIdx.x; i < nw; i += blockDim.x) argmin_merge(x, idx, cw[i], ci[i]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double ox = __shfl_xor_sync(0xffffffffu, x, off);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, off);
        argmin_merge(x, idx, ox, oi);
    }
    if ((threadIdx.x & 31) == 0) { sw[threadIdx.x >> 5] = x; si[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 32; k++) argmin_merge(x, idx, sw[k], si[k]);
        if (idx == 0x7fffffff) idx = 0;  // all weights NaN: std::min_element keeps the first
        out[0] = xv[idx]; out[1] = xv[pp + idx]; out[2] = xv[2 * pp + idx];
        out[3] = (double)idx;
    }
}

__global__ void k_get_features(const double* __restrict__ xf, const double* __restrict__ pf, size_t pp, int p,
                               int nf, double* __restrict__ XF, double* __restrict__ PF) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    XF[2 * f] = xf[(size_t)(2 * f) * pp + p];
    XF[2 * f + 1] = xf[(size_t)(2 * f + 1) * pp + p];
    const double xx = pf[(size_t)(3 * f) * pp + p], xy = pf[(size_t)(3 * f + 1) * pp + p],
                 yy = pf[(size_t)(3 * f + 2) * pp + p];
    PF[4 * f] = xx; PF[4 * f + 1] = xy; PF[4 * f + 2] = xy; PF[4 * f + 3] = yy;
}
// [p][k] (AoS, host convention) <-> [k][pp] (SoA)
__global__ void __launch_bounds__(256) k_aos_to_soa(const double* __restrict__ aos, double* __restrict__ soa,
                                                    size_t pp, int np, int k) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)np * k) return;
    const int p = (int)(idx / k), j = (int)(idx % k);
    soa[(size_t)j * pp + p] = aos[idx];
}
__global__ void __launch_bounds__(256) k_soa_to_aos(const double* __restrict__ soa, double* __restrict__ aos,
                                                    size_t pp, int np, int k) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)np * k) return;
    const int p = (int)(idx / k), j = (int)(idx % k);
    aos[idx] = soa[(size_t)j * pp + p];
}

static int check_pf(const cslam_pf* h) {
    CSLAM_REQUIRE(h != nullptr, CSLAM_ERR_BAD_ARG, "null handle");
    CSLAM_CUDA(cudaSetDevice(h->device));
    return CSLAM_OK;
}
static inline unsigned nblk(size_t n, int b) { return (unsigned)((n + b - 1) / b); }

static int fill_obs(PfObs& ob, const double* Z, const int32_t* idf, int m, const double R[4]) {
    memset(&ob, 0, sizeof(ob));
    memcpy(ob.z, Z, sizeof(double) * 2 * m);
    if (idf) memcpy(ob.idf, idf, sizeof(int) * m);
    ob.m = m;
    ob.R[0][0] = R[0]; ob.R[1][0] = R[1]; ob.R[0][1] = R[2]; ob.R[1][1] = R[3];
    return CSLAM_OK;
}

// host->device staging of a caller array (pageable): synchronous copy into d_in
static int stage_in(cslam_pf* h, const double* src, size_t count, int on_device, const double** out) {
    if (on_device) {
        *out = src;
        return CSLAM_OK;
    }
    CSLAM_CUDA(cudaMemcpyAsync(h->d_in, src, count * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    *out = h->d_in;
    return CSLAM_OK;
}

// Canonical hierarchical scan of f(in).  Local levels (whole 32-groups inside this rank) go to
// h->scan[l]; from h->gather_level on, the group totals of ALL ranks are all-gathered in rank order
// and the remaining levels are computed redundantly on every rank in h->gscan[l] — exactly the
// arrays a single GPU would compute, so every sum is bit-identical for any rank count.
static int run_scan(cslam_pf* h, const double* in, int mode, const double* div) {
    const double* src = in;
    size_t len = h->np;
    const int L = h->gather_level < 0 ? h->levels : h->gather_level;
    for (int l = 0; l < L; l++) {
        double* totals = (l + 1 < h->levels) ? h->scan[l + 1] : h->d_small + 7;
        count_launch();
        k_scan_level<<<nblk(len, 256), 256, 0, h->stream>>>(src, len, h->scan[l], totals, l == 0 ? mode : 0, div);
        if (l + 1 < h->levels) src = h->scan[l + 1];  // level l+1's INPUT is the totals array; it is scanned in place
        len = (len + 31) / 32;
    }
    if (L < h->levels) {
        const NcclApi* api = nccl_api();
        if (!api) return CSLAM_ERR_NCCL;
        CSLAM_NCCL(api->AllGather(h->scan[L], h->gscan[L], len, ncclDouble, h->comm, h->stream));
        size_t glen = len * (size_t)h->world;
        for (int l = L; l < h->levels; l++) {
            double* totals = (l + 1 < h->levels) ? h->gscan[l + 1] : h->d_small + 7;
            count_launch();
            k_scan_level<<<nblk(glen, 256), 256, 0, h->stream>>>(h->gscan[l], glen, h->gscan[l], totals, 0, div);
            glen = (glen + 31) / 32;
        }
    }
    CSLAM_CUDA(cudaGetLastError());
    return CSLAM_OK;
}
static ScanPtrs scan_ptrs(const cslam_pf* h) {
    ScanPtrs sp;
    const int L = h->gather_level < 0 ? h->levels : h->gather_level;
    for (int l = 0; l < 5; l++) sp.s[l] = (l >= L) ? h->gscan[l] : h->scan[l];
    sp.levels = h->levels;
    sp.glevel = L;
    sp.base = (long long)h->rank * h->np;
    return sp;
}
static const double* scan_top(const cslam_pf* h, size_t* len) {
    const int t = h->levels - 1;
    const bool g = h->gather_level >= 0 && t >= h->gather_level;
    *len = g ? h->gscan_len[t] : h->scan_len[t];
    return g ? h->gscan[t] : h->scan[t];
}

}  // namespace cslam

using namespace cslam;

extern "C" {

static int pf_create_common(cslam_pf_t** out, int num_particles, int capacity_landmarks, int device, unsigned flags,
                            int rank, int world, const void* nccl_id) {
    CSLAM_REQUIRE(out != nullptr, CSLAM_ERR_BAD_ARG, "out is null");
    CSLAM_REQUIRE(num_particles >= 1 && capacity_landmarks >= 0, CSLAM_ERR_BAD_ARG, "bad sizes");
    CSLAM_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world, CSLAM_ERR_BAD_ARG, "bad rank/world");
    CSLAM_REQUIRE(world == 1 || (nccl_id != nullptr && num_particles % 32 == 0), CSLAM_ERR_BAD_ARG,
                  "sharded PF needs an NCCL id and a local particle count that is a multiple of 32");
    CSLAM_REQUIRE((long long)num_particles * world < (1LL << 31), CSLAM_ERR_UNSUPPORTED, "too many particles");
    *out = nullptr;
    int count = 0;
    CSLAM_CUDA(cudaGetDeviceCount(&count));
    CSLAM_REQUIRE(device >= 0 && device < count, CSLAM_ERR_CUDA, "no such CUDA device (no CPU fallback exists)");
    CSLAM_CUDA(cudaSetDevice(device));
    cslam_pf* h = new (std::nothrow) cslam_pf();
    CSLAM_REQUIRE(h != nullptr, CSLAM_ERR_BAD_ARG, "out of host memory");
    h->device = device;
    h->flags = flags;
    h->np = num_particles;
    h->pp = ((size_t)num_particles + 31) / 32 * 32;
    h->nf_cap = capacity_landmarks;
    h->rank = rank;
    h->world = world;
    h->np_global = (long long)num_particles * world;
    auto fail = [&](int code) {
        cslam_pf_destroy(h);
        return code;
    };
#define TRY(call)                                                                           \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            set_last_error("cslam_pf_create: %s -> %s", #call, cudaGetErrorString(e__));    \
            return fail(CSLAM_ERR_CUDA);                                                    \
        }                                                                                   \
    } while (0)
    TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
    const size_t pp = h->pp;
    for (int b = 0; b < 2; b++) {
        TRY(cudaMalloc(&h->buf[b].w, pp * sizeof(double)));
        TRY(cudaMalloc(&h->buf[b].xv, 3 * pp * sizeof(double)));
        TRY(cudaMalloc(&h->buf[b].pv, 9 * pp * sizeof(double)));
        TRY(cudaMalloc(&h->buf[b].xf, std::max<size_t>(1, 2 * (size_t)h->nf_cap) * pp * sizeof(double)));
        TRY(cudaMalloc(&h->buf[b].pf, std::max<size_t>(1, 3 * (size_t)h->nf_cap) * pp * sizeof(double)));
        TRY(cudaMemsetAsync(h->buf[b].w, 0, pp * sizeof(double), h->stream));
        TRY(cudaMemsetAsync(h->buf[b].xv, 0, 3 * pp * sizeof(double), h->stream));
        TRY(cudaMemsetAsync(h->buf[b].pv, 0, 9 * pp * sizeof(double), h->stream));
        h->peer[b][rank] = h->buf[b];
    }
    // scan hierarchy: the level structure is that of the GLOBAL particle count (identical for any
    // world size); a level is local while this rank's share of its input is whole 32-groups.
    {
        size_t glen[6];
        int levels = 0;
        size_t len = (size_t)h->np_global;
        while (true) {
            glen[levels++] = len;
            if (len <= 32) break;
            len = (len + 31) / 32;
            if (levels >= 5) break;
        }
        if (glen[levels - 1] > 32) {
            set_last_error("cslam_pf_create: too many particles for a 5-level radix-32 scan");
            return fail(CSLAM_ERR_UNSUPPORTED);
        }
        h->levels = levels;
        size_t llen = h->np;
        int L = 0;
        if (world == 1) {
            L = levels;
        } else {
            while (L < levels && llen % 32 == 0) { L++; llen /= 32; }
            h->gather_level = L;
        }
        llen = h->np;
        for (int l = 0; l <= L && l < levels; l++) {  // scan[L] holds the local totals that get all-gathered
            h->scan_len[l] = (world == 1) ? glen[l] : llen;
            TRY(cudaMalloc(&h->scan[l], (h->scan_len[l] + 32) * sizeof(double)));
            llen = (llen + 31) / 32;
        }
        for (int l = L; l < levels; l++) {
            h->gscan_len[l] = glen[l];
            TRY(cudaMalloc(&h->gscan[l], (glen[l] + 32) * sizeof(double)));
        }
    }
    TRY(cudaMalloc(&h->comb, pp * sizeof(double)));
    TRY(cudaMalloc(&h->wn, pp * sizeof(double)));
    TRY(cudaMalloc(&h->keep, pp * sizeof(int)));
    TRY(cudaMalloc(&h->d_in, 3 * pp * sizeof(double)));
    TRY(cudaMalloc(&h->d_small, 16 * sizeof(double)));
    TRY(cudaMalloc(&h->d_ismall, 8 * sizeof(int)));
    TRY(cudaMemsetAsync(h->d_ismall, 0, 8 * sizeof(int), h->stream));
    h->pinned_bytes = 1 << 16;
    TRY(cudaMallocHost(&h->pinned, h->pinned_bytes));
    // PF.cpp:319-341: w = 1/P
    count_launch();
    k_fill<<<nblk(h->np, 256), 256, 0, h->stream>>>(h->buf[0].w, h->np, 1.0 / (double)h->np_global);
    // PF.cpp:581-587: DI(0) = k/2 ; DI(i) = DI(i-1) + k  — a sequential chain over ALL slots, built once
    // on the host; every rank keeps the slice of its own slots
    {
        std::vector<double> comb(h->np);
        const double k = 1.0 / (double)h->np_global;
        double di = k / 2.0;
        const long long first = (long long)rank * h->np;
        for (long long i = 0; i < first + h->np; i++) {
            if (i > 0) di = di + k;
            if (i >= first) comb[(size_t)(i - first)] = di;
        }
        TRY(cudaMemcpyAsync(h->comb, comb.data(), (size_t)h->np * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        TRY(cudaStreamSynchronize(h->stream));
    }
    if (world > 1) {
        TRY(cudaMalloc(&h->cum_global, (size_t)h->np_global * sizeof(double)));
        const NcclApi* api = nccl_api();
        if (!api) return fail(CSLAM_ERR_NCCL);
        ncclUniqueId id;
        memcpy(&id, nccl_id, sizeof(id));
        ncclResult_t r = api->CommInitRank(&h->comm, world, id, rank);
        if (r != ncclSuccess) {
            set_last_error("ncclCommInitRank -> %s", api->GetErrorString(r));
            return fail(CSLAM_ERR_NCCL);
        }
    }
    TRY(cudaStreamSynchronize(h->stream));
#undef TRY
    *out = h;
    return CSLAM_OK;
}

int cslam_pf_create(cslam_pf_t** out, int num_particles, int capacity_landmarks, int device, unsigned flags) {
    CSLAM_NVTX_RANGE();
    return pf_create_common(out, num_particles, capacity_landmarks, device, flags, 0, 1, nullptr);
}

int cslam_pf_create_sharded(cslam_pf_t** out, int num_particles_local, int capacity_landmarks, int device,
                            unsigned flags, int rank, int world, const void* nccl_unique_id) {
    CSLAM_NVTX_RANGE();
    return pf_create_common(out, num_particles_local, capacity_landmarks, device, flags, rank, world, nccl_unique_id);
}

// 10 cudaIpcMemHandle_t (64 bytes each): {w, xv, pv, xf, pf} of both ping-pong buffers
int cslam_pf_ipc_export(cslam_pf_t* h, void* out640) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_pf(h)) return rc;
    CSLAM_REQUIRE(out640 != nullptr, CSLAM_ERR_BAD_ARG, "null");
    cudaIpcMemHandle_t* hs = static_cast<cudaIpcMemHandle_t*>(out640);
    for (int b = 0; b < 2; b++) {
        CSLAM_CUDA(cudaIpcGetMemHandle(&hs[5 * b + 0], h->buf[b].w));
        CSLAM_CUDA(cudaIpcGetMemHandle(&hs[5 * b + 1], h->buf[b].xv));
        CSLAM_CUDA(cudaIpcGetMemHandle(&hs[5 * b + 2], h->buf[b].pv));
        CSLAM_CUDA(cudaIpcGetMemHandle(&hs[5 * b + 3], h->buf[b].xf));
        CSLAM_CUDA(cudaIpcGetMemHandle(&hs[5 * b + 4], h->buf[b].pf));
    }
    return CSLAM_OK;
}
// all: world x 640 bytes in rank order (every rank's export).  Maps the peers' buffers (NVLink P2P).
int cslam_pf_ipc_import(cslam_pf_t* h, const void* all, int world) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_pf(h)) return rc;
    CSLAM_REQUIRE(all != nullptr && world == h->world, CSLAM_ERR_BAD_ARG, "bad argument");
    const cudaIpcMemHandle_t* hs = static_cast<const cudaIpcMemHandle_t*>(all);
    for (int r = 0; r < world; r++) {
        if (r == h->rank) continue;
        for (int b = 0; b < 2; b++) {
            void* p[5];
            for (int k = 0; k < 5; k++)
                CSLAM_CUDA(cudaIpcOpenMemHandle(&p[k], hs[10 * r + 5 * b + k], cudaIpcMemLazyEnablePeerAccess));
            h->peer[b][r].w = static_cast<double*>(p[0]);
            h->peer[b][r].xv = static_cast<double*>(p[1]);
            h->peer[b][r].pv = static_cast<double*>(p[2]);
            h->peer[b][r].xf = static_cast<double*>(p[3]);
            h->peer[b][r].pf = static_cast<double*>(p[4]);
        }
    }
    h->peers_ready = true;
    return CSLAM_OK;
}

int cslam_pf_destroy(cslam_pf_t* h) {
    CSLAM_NVTX_RANGE();
    if (!h) return CSLAM_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->peers_ready) {
        for (int r = 0; r < h->world; r++) {
            if (r == h->rank) continue;
            for (int b = 0; b < 2; b++) {
                cudaIpcCloseMemHandle(h->peer[b][r].w); cudaIpcCloseMemHandle(h->peer[b][r].xv);
                cudaIpcCloseMemHandle(h->peer[b][r].pv); cudaIpcCloseMemHandle(h->peer[b][r].xf);
                cudaIpcCloseMemHandle(h->peer[b][r].pf);
            }
        }
    }
    if (h->comm) {
        const NcclApi* api = nccl_api();
        if (api) api->CommDestroy(h->comm);
    }
    for (int b = 0; b < 2; b++) {
        cudaFree(h->buf[b].w); cudaFree(h->buf[b].xv); cudaFree(h->buf[b].pv);
        cudaFree(h->buf[b].xf); cudaFree(h->buf[b].pf);
    }
    for (int l = 0; l < 5; l++) { cudaFree(h->scan[l]); cudaFree(h->gscan[l]); }
    cudaFree(h->cum_global);
    cudaFree(h->d_peer_tab);
    cudaFree(h->acc_dev);
    cudaFree(h->comb); cudaFree(h->wn); cudaFree(h->keep); cudaFree(h->d_in);
    cudaFree(h->d_small); cudaFree(h->d_ismall);
    if (h->pinned) cudaFreeHost(h->pinned);
    for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return CSLAM_OK;
}

int cslam_pf_set_stream(cslam_pf_t* h, void* cuda_stream) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_pf(h)) return rc;
    CSLAM_CUDA(cudaStreamSynchronize(h->stream));
    if (h->own_stream) cudaStreamDestroy(h->stream);
    h->stream = static_cast<cudaStream_t>(cuda_stream);
    h->own_stream = false;
    return CSLAM_OK;
}

int cslam_pf_sync(cslam_pf_t* h, int* skipped_updates) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_pf(h)) return rc;
    if (skipped_updates) {
        CSLAM_CUDA(cudaMemcpyAsync(h->pinned, h->d_ismall + 2, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CSLAM_CUDA(cudaStreamSynchronize(h->stream));
        *skipped_updates = *static_cast<int*>(h->pinned);
    } else {
        CSLAM_CUDA(cudaStreamSynchronize(h->stream));
    }
    return CSLAM_OK;
}

int cslam_pf_num_particles(const cslam_pf_t* h) { return h ? h->np : -1; }
int cslam_pf_num_features(const cslam_pf_t* h) { return h ? h->nf : -1; }

int cslam_pf_predict(cslam_pf_t* h, double v, double swa, const double Q[4], double wb, double dt) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_pf(h)) return rc;
    CSLAM_REQUIRE(Q != nullptr, CSLAM_ERR_BAD_ARG, "Q is null");
    PfBuf& b = h->buf[h->cur];
    count_launch();
    k_pf_predict<<<nblk(h->np, 256), 256, 0, h->stream>>>(b.xv, b.pv, h->pp, h->np, v, swa, Q[0], Q[2], Q[1], Q[3],
                                                          wb, dt);
    CSLAM_CUDA(cudaGetLastError());
    return CSLAM_OK;
}

int cslam_pf_observe_heading(cslam_pf_t* h, double phi, int use_heading) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_pf(h)) return rc;
    if (!use_heading) return CSLAM_OK;
    const double sigma = 0.01F * kPi / 180.0F;  // PF.cpp:391
    PfBuf& b = h->buf[h->cur];
    count_launch();
    k_pf_heading<<<nblk(h->np, 256), 256, 0, h->stream>>>(b.xv, b.pv, h->pp, h->np, phi, sigma * sigma);
    CSLAM_CUDA(cudaGetLastError());
    return CSLAM_OK;
}

int cslam_pf_control_steps(cslam_pf_t* h, int k, const double* v, const double* swa, const double* phi,
                           int use_heading, const double Q[4], double wb, double dt) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_pf(h)) return rc;
    CSLAM_REQUIRE(k >= 0, CSLAM_ERR_BAD_ARG, "k < 0");
    if (k == 0) return CSLAM_OK;
    CSLAM_REQUIRE(v && swa && Q && (phi || !use_heading), CSLAM_ERR_BAD_ARG, "null argument");
    const double sigma = 0.01F * kPi / 180.0F;  // PF.cpp:391
    PfBuf& b = h->buf[h->cur];
    for (int base = 0; base < k; base += kPfMaxControlSteps) {
        PfControlPack cp;
        memset(&cp, 0, sizeof(cp));
        cp.k = std::min(kPfMaxControlSteps, k - base);
        for (int i = 0; i < cp.k; i++) {
            cp.v[i] = v[base + i];
            cp.swa[i] = swa[base + i];
            cp.phi[i] = phi ? phi[base + i] : 0.0;
        }
        cp.use_heading = use_heading ? 1 : 0;
        cp.q00 = Q[0]; cp.q01 = Q[2]; cp.q10 = Q[1]; cp.q11 = Q[3];
        cp.wb = wb; cp.dt = dt; cp.rh = sigma * sigma;
        count_launch();
        k_pf_control_steps<<<nblk(h->np, 256), 256, 0, h->stream>>>(b.xv, b.pv, h->pp, h->np, cp);
        CSLAM_CUDA(cudaGetLastError());
    }
    return CSLAM_OK;
}

static int check_obs(cslam_pf* h, const double* Z, const int32_t* idf, int m, const double R[4], bool need_idf) {
    CSLAM_REQUIRE(m >= 0 && m <= CSLAM_MAX_OBS, CSLAM_ERR_BAD_ARG, "m out of range (0..CSLAM_MAX_OBS)");
    if (m == 0) return CSLAM_OK;
    CSLAM_REQUIRE(Z && R && (!need_idf || idf), CSLAM_ERR_BAD_ARG, "null argument");
    if (need_idf) {
        for (int i = 0; i < m; i++) {
            CSLAM_REQUIRE(idf[i] >= 1 && idf[i] <= h->nf, CSLAM_ERR_BAD_ARG, "idf out of range (1-based)");
            for (int j = 0; j < i; j++)
                CSLAM_REQUIRE(idf[j] != idf[i], CSLAM_ERR_UNSUPPORTED, "duplicate landmark id within one scan");
        }
    }
    return CSLAM_OK;
}

int cslam_pf_sample_proposal(cslam_pf_t* h, const double* Z, const int32_t* idf, int m, const double R[4],
                             const double* xi, int xi_on_device) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_pf(h)) return rc;
    if (int rc = check_obs(h, Z, idf, m, R, true)) return rc;
    CSLAM_REQUIRE(xi != nullptr, CSLAM_ERR_BAD_ARG, "xi is null");
    PfObs ob;
    if (m > 0) fill_obs(ob, Z, idf, m, R); else { memset(&ob, 0, sizeof(ob)); }
    const double* dxi = nullptr;
    if (int rc = stage_in(h, xi, (size_t)3 * h->np, xi_on_device, &dxi)) return rc;
    PfBuf& b = h->buf[h->cur];
    count_launch();
    k_pf_sample_proposal<<<nblk(h->np, 128), 128, 0, h->stream>>>(b.w, b.xv, b.pv, b.xf, b.pf, h->pp, h->np, ob, dxi,
                                                                  h->flags, h->d_ismall + 2);
    CSLAM_CUDA(cudaGetLastError());
    return CSLAM_OK;
}

int cslam_pf_feature_update(cslam_pf_t* h, const double* Z, const int32_t* idf, int m, const double R[4]) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_pf(h)) return rc;
    if (int rc = check_obs(h, Z, idf, m, R, true)) return rc;
    if (m == 0) return CSLAM_OK;
    PfObs ob;
    fill_obs(ob, Z, idf, m, R);
    PfBuf& b = h->buf[h->cur];
    count_launch();
    k_pf_feature_update<<<nblk(h->np, 128), 128, 0, h->stream>>>(b.xv, b.xf, b.pf, h->pp, h->np, ob, h->flags,
                                                                 h->d_ismall + 2);
    CSLAM_CUDA(cudaGetLastError());
    return CSLAM_OK;
}

int cslam_pf_add_features(cslam_pf_t* h, const double* Z, int m, const double R[4]) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_pf(h)) return rc;
    if (int rc = check_obs(h, Z, nullptr, m, R, false)) return rc;
    if (m == 0) return CSLAM_OK;
    CSLAM_REQUIRE(h->nf + m <= h->nf_cap, CSLAM_ERR_CAPACITY, "landmark capacity exceeded");
    PfObs ob;
    fill_obs(ob, Z, nullptr, m, R);
    PfBuf& b = h->buf[h->cur];
    count_launch();
    k_pf_add_features<<<nblk(h->np, 256), 256, 0, h->stream>>>(b.xv, b.xf, b.pf, h->pp, h->np, h->nf, ob);
    CSLAM_CUDA(cudaGetLastError());
    h->nf += m;
    return CSLAM_OK;
}

int cslam_pf_sample_pose(cslam_pf_t* h, const double* xi, int xi_on_device) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_pf(h)) return rc;
    CSLAM_REQUIRE(xi != nullptr, CSLAM_ERR_BAD_ARG, "xi is null");
    const double* dxi = nullptr;
    if (int rc = stage_in(h, xi, (size_t)3 * h->np, xi_on_device, &dxi)) return rc;
    PfBuf& b = h->buf[h->cur];
    count_launch();
    k_pf_sample_pose<<<nblk(h->np, 256), 256, 0, h->stream>>>(b.xv, b.pv, h->pp, h->np, dxi, h->d_ismall + 2);
    CSLAM_CUDA(cudaGetLastError());
    return CSLAM_OK;
}

int cslam_pf_resample(cslam_pf_t* h, const double* u, int u_on_device, double num_effective, int resample_on,
                      int32_t* keep, double* neff, int* resampled) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_pf(h)) return rc;
    CSLAM_REQUIRE(u != nullptr, CSLAM_ERR_BAD_ARG, "u is null");
    const int np = h->np;
    PfBuf& b = h->buf[h->cur];
    const double* du = nullptr;
    if (int rc = stage_in(h, u, (size_t)np, u_on_device, &du)) return rc;
    double* S = h->d_small;  // S[0] = sum w, S[1] = sum W, S[2] = sum W^2
    const bool intended = (h->flags & CSLAM_FLAG_Q10_SEARCH) != 0;
    CSLAM_REQUIRE(h->world == 1 || (intended && h->peers_ready), CSLAM_ERR_UNSUPPORTED,
                  "multi-GPU resampling needs CSLAM_FLAG_Q10_SEARCH and cslam_pf_ipc_import");
    const ScanPtrs sp = scan_ptrs(h);
    size_t top_len = 0;
    const double* top = scan_top(h, &top_len);
    if (intended) {
        // ws = sum w ; particles.w /= ws (PF.cpp:482-487)
        if (int rc = run_scan(h, b.w, 0, nullptr)) return rc;
        count_launch();
        k_scan_total<<<1, 32, 0, h->stream>>>(top, top_len, S + 0);
        count_launch();
        k_divide<<<nblk(np, 256), 256, 0, h->stream>>>(b.w, np, S + 0);
        // stratifiedResample: W /= W.sum() (PF.cpp:548) ; neff = 1 / sum W^2 (:550-554)
        if (int rc = run_scan(h, b.w, 0, nullptr)) return rc;
        count_launch();
        k_scan_total<<<1, 32, 0, h->stream>>>(top, top_len, S + 1);
        if (int rc = run_scan(h, b.w, 2, S + 1)) return rc;
        count_launch();
        k_scan_total<<<1, 32, 0, h->stream>>>(top, top_len, S + 2);
        // cumulative sum of W (PF.cpp:559-564) in the canonical order
        if (int rc = run_scan(h, b.w, 1, S + 1)) return rc;
        count_launch();
        k_scan_combine<<<nblk(np, 256), 256, 0, h->stream>>>(sp, np, h->wn);
        const double* cum = h->wn;
        if (h->world > 1) {  // every rank searches the cumulative weights of ALL particles
            const NcclApi* api = nccl_api();
            if (!api) return CSLAM_ERR_NCCL;
            CSLAM_NCCL(api->AllGather(h->wn, h->cum_global, (size_t)np, ncclDouble, h->comm, h->stream));
            cum = h->cum_global;
        }
        count_launch();
        k_resample_search<<<nblk(np, 256), 256, 0, h->stream>>>(cum, (int)h->np_global, h->comb, du, np, h->keep);
    } else {
        count_launch();
        k_seq_sum<<<1, 32, 0, h->stream>>>(b.w, np, 0, nullptr, S + 0);
        count_launch();
        k_divide<<<nblk(np, 256), 256, 0, h->stream>>>(b.w, np, S + 0);
        count_launch();
        k_seq_sum<<<1, 32, 0, h->stream>>>(b.w, np, 0, nullptr, S + 1);
        count_launch();
        k_seq_sum<<<1, 32, 0, h->stream>>>(b.w, np, 2, S + 1, S + 2);
        count_launch();
        k_seq_cumsum<<<1, 32, 0, h->stream>>>(b.w, np, S + 1, h->wn);
        CSLAM_CUDA(cudaMemsetAsync(h->d_ismall, 0x7f, sizeof(int), h->stream));  // 0x7f7f7f7f = "no hit"
        count_launch();
        k_resample_first_hit<<<nblk(np, 256), 256, 0, h->stream>>>(h->wn, h->comb, du, np, h->d_ismall);
        count_launch();
        k_fill_keep<<<nblk(np, 256), 256, 0, h->stream>>>(h->keep, np, h->d_ismall);
    }
    CSLAM_CUDA(cudaGetLastError());
    // neff decides whether the gather runs (PF.cpp:490: neff < numEffective).  With numEffective = +inf
    // ("resample every step") the decision does not depend on neff: no read-back, the call never waits for
    // the GPU unless the caller asks for neff or the indices.
    const bool always = resample_on && isinf(num_effective) && num_effective > 0;
    double ne = 0.0;
    if (!always || neff) {
        CSLAM_CUDA(cudaMemcpyAsync(h->pinned, S + 2, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CSLAM_CUDA(cudaStreamSynchronize(h->stream));
        ne = 1.0 / *static_cast<double*>(h->pinned);
    }
    if (neff) *neff = ne;
    if (keep) {
        CSLAM_CUDA(cudaMemcpyAsync(keep, h->keep, (size_t)np * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CSLAM_CUDA(cudaStreamSynchronize(h->stream));
    }
    const bool doit = always || ((ne < num_effective) && resample_on);
    if (resampled) *resampled = doit ? 1 : 0;
    if (doit) {
        PfBuf& d = h->buf[h->cur ^ 1];
        const unsigned gx = nblk(((size_t)np + 1) / 2, 256);
        if (!h->d_peer_tab) {  // device table of every rank's base pointers: [buffer][xv,pv,xf,pf][rank]
            CSLAM_CUDA(cudaMalloc(&h->d_peer_tab, 2 * 4 * 8 * sizeof(double*)));
            const double* tab[2][4][8] = {};
            for (int bb = 0; bb < 2; bb++)
                for (int r = 0; r < h->world; r++) {
                    tab[bb][0][r] = h->peer[bb][r].xv; tab[bb][1][r] = h->peer[bb][r].pv;
                    tab[bb][2][r] = h->peer[bb][r].xf; tab[bb][3][r] = h->peer[bb][r].pf;
                }
            CSLAM_CUDA(cudaMemcpyAsync(h->d_peer_tab, tab, sizeof(tab), cudaMemcpyHostToDevice, h->stream));
            CSLAM_CUDA(cudaStreamSynchronize(h->stream));
        }
        auto gather = [&](double* PfBuf::*field, int slot, int rows) {
            if (rows <= 0) return;
            const unsigned gy = (unsigned)std::min(rows, 65535);
            count_launch();
            static const bool force_peer = getenv("CSLAM_PF_FORCE_PEER_KERNEL") != nullptr;  // development switch
            if (h->world == 1 && !force_peer) {
                k_gather_rows<<<dim3(gx, gy), 256, 0, h->stream>>>(b.*field, d.*field, h->pp, np, rows, h->keep);
            } else {
                k_gather_rows_peer<<<dim3(gx, gy), 256, 0, h->stream>>>(h->d_peer_tab + (h->cur * 4 + slot) * 8,
                                                                        d.*field, h->pp, np, rows, h->keep);
            }
        };
        const bool prof = h->prof && h->prof_used + 2 <= (int)h->prof_ev.size();
        if (prof) cudaEventRecord(h->prof_ev[h->prof_used], h->stream);
        gather(&PfBuf::xv, 0, 3);
        gather(&PfBuf::pv, 1, 9);
        gather(&PfBuf::xf, 2, 2 * h->nf);
        gather(&PfBuf::pf, 3, 3 * h->nf);
        if (prof) {
            cudaEventRecord(h->prof_ev[h->prof_used + 1], h->stream);
            h->prof_used += 2;
        }
        count_launch();
        k_fill<<<nblk(np, 256), 256, 0, h->stream>>>(d.w, np, 1.0 / (double)h->np_global);  // PF.cpp:495
        CSLAM_CUDA(cudaGetLastError());
        h->cur ^= 1;
    }
    return CSLAM_OK;
}

// w[p] *= factor[p] (device or host vector of np doubles): importance-weight modulation by an external
// likelihood term; bench.py uses it to study resampling under realistic / adversarial weight spreads.
__global__ void __launch_bounds__(256) k_scale_w(double* __restrict__ w, const double* __restrict__ f, int np) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < np) w[i] = w[i] * f[i];
}
int cslam_pf_scale_weights(cslam_pf_t* h, const double* factor, int on_device) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_pf(h)) return rc;
    CSLAM_REQUIRE(factor != nullptr, CSLAM_ERR_BAD_ARG, "factor is null");
    const double* df = nullptr;
    if (int rc = stage_in(h, factor, (size_t)h->np, on_device, &df)) return rc;
    count_launch();
    k_scale_w<<<nblk(h->np, 256), 256, 0, h->stream>>>(h->buf[h->cur].w, df, h->np);
    CSLAM_CUDA(cudaGetLastError());
    return CSLAM_OK;
}

int cslam_pf_profile_begin(cslam_pf_t* h, int max_resamples) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_pf(h)) return rc;
    CSLAM_REQUIRE(max_resamples > 0 && max_resamples <= (1 << 16), CSLAM_ERR_BAD_ARG, "out of range");
    while ((int)h->prof_ev.size() < 2 * max_resamples) {
        cudaEvent_t e;
        CSLAM_CUDA(cudaEventCreate(&e));
        h->prof_ev.push_back(e);
    }
    h->prof_used = 0;
    h->prof = true;
    return CSLAM_OK;
}
int cslam_pf_profile_end(cslam_pf_t* h, double* ms, int* resamples, double* bytes) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_pf(h)) return rc;
    h->prof = false;
    CSLAM_CUDA(cudaStreamSynchronize(h->stream));
    double total = 0.0;
    for (int i = 0; i + 1 < h->prof_used; i += 2) {
        float t = 0.f;
        CSLAM_CUDA(cudaEventElapsedTime(&t, h->prof_ev[i], h->prof_ev[i + 1]));
        total += t;
    }
    if (ms) *ms = total;
    if (resamples) *resamples = h->prof_used / 2;
    // algorithmic bytes of one gather-copy: every SoA row of every local particle read once, written once
    if (bytes) *bytes = (double)(h->prof_used / 2) * 2.0 * 8.0 * (12.0 + 5.0 * h->nf) * (double)h->np;
    return CSLAM_OK;
}

// Checkpoint of the particle set (the reference keeps std::vector<Particle_t> in the driver and has no
// persistence): header {magic, version, flags, np, nf} + every SoA row of the current
// buffer (w, xv[3], pv[9], xf[2*nf], pf[3*nf]), np doubles each.  Single-GPU handles.
namespace {
struct PfCkptHeader {
    char magic[8];
    uint32_t version;
    uint32_t flags;
    int32_t np;
    int32_t nf;
};
const char kPfCkptMagic[8] = {'C', 'S', 'L', 'A', 'M', 'P', 'F', '1'};
struct PfRows {
    double* base;
    int rows;
};
}  // namespace

static int pf_ckpt_io(cslam_pf* h, FILE* f, bool save) {
    PfBuf& b = h->buf[h->cur];
    const PfRows groups[5] = {{b.w, 1}, {b.xv, 3}, {b.pv, 9}, {b.xf, 2 * h->nf}, {b.pf, 3 * h->nf}};
    const size_t row_bytes = (size_t)h->np * sizeof(double);
    const int slab = (int)std::max<size_t>(1, (size_t)(64u << 20) / row_bytes);  // rows per ~64 MB transfer
    std::vector<double> host((size_t)slab * h->np);
    for (const PfRows& g : groups) {
        for (int r0 = 0; r0 < g.rows; r0 += slab) {
            const int nr = std::min(slab, g.rows - r0);
            if (save) {
                CSLAM_CUDA(cudaMemcpy2D(host.data(), row_bytes, g.base + (size_t)r0 * h->pp, h->pp * sizeof(double),
                                        row_bytes, nr, cudaMemcpyDeviceToHost));
                CSLAM_REQUIRE(fwrite(host.data(), row_bytes, nr, f) == (size_t)nr, CSLAM_ERR_BAD_ARG,
                              "short write to the checkpoint file");
            } else {
                CSLAM_REQUIRE(fread(host.data(), row_bytes, nr, f) == (size_t)nr, CSLAM_ERR_BAD_ARG,
                              "truncated checkpoint file");
                CSLAM_CUDA(cudaMemcpy2D(g.base + (size_t)r0 * h->pp, h->pp * sizeof(double), host.data(), row_bytes,
                                        row_bytes, nr, cudaMemcpyHostToDevice));
            }
        }
    }
    return CSLAM_OK;
}

int cslam_pf_save(cslam_pf_t* h, const char* path) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_pf(h)) return rc;
    CSLAM_REQUIRE(path != nullptr, CSLAM_ERR_BAD_ARG, "path is null");
    CSLAM_REQUIRE(h->world == 1, CSLAM_ERR_UNSUPPORTED, "checkpoints are single-GPU (sharded: save per rank)");
    CSLAM_CUDA(cudaStreamSynchronize(h->stream));
    FILE* f = fopen(path, "wb");
    CSLAM_REQUIRE(f != nullptr, CSLAM_ERR_BAD_ARG, "cannot open checkpoint file for writing");
    PfCkptHeader hd;
    memcpy(hd.magic, kPfCkptMagic, 8);
    hd.version = 1;
    hd.flags = h->flags;
    hd.np = h->np;
    hd.nf = h->nf;
    int rc = fwrite(&hd, sizeof(hd), 1, f) == 1 ? CSLAM_OK : CSLAM_ERR_BAD_ARG;
    if (rc == CSLAM_OK) rc = pf_ckpt_io(h, f, true);
    if (fclose(f) != 0 && rc == CSLAM_OK) rc = CSLAM_ERR_BAD_ARG;
    return rc;
}

int cslam_pf_load(cslam_pf_t* h, const char* path) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_pf(h)) return rc;
    CSLAM_REQUIRE(path != nullptr, CSLAM_ERR_BAD_ARG, "path is null");
    CSLAM_REQUIRE(h->world == 1, CSLAM_ERR_UNSUPPORTED, "checkpoints are single-GPU (sharded: load per rank)");
    FILE* f = fopen(path, "rb");
    CSLAM_REQUIRE(f != nullptr, CSLAM_ERR_BAD_ARG, "cannot open checkpoint file");
    PfCkptHeader hd;
    const bool ok = fread(&hd, sizeof(hd), 1, f) == 1 && memcmp(hd.magic, kPfCkptMagic, 8) == 0 && hd.version == 1 &&
                    hd.np == h->np && hd.nf >= 0 && hd.nf <= h->nf_cap;
    if (!ok) {
        fclose(f);
        set_last_error("cslam_pf_load: not a particle checkpoint of this library, or particle count / landmark "
                       "capacity of the handle do not match");
        return CSLAM_ERR_BAD_ARG;
    }
    CSLAM_CUDA(cudaStreamSynchronize(h->stream));
    h->nf = hd.nf;
    const int rc = pf_ckpt_io(h, f, false);
    fclose(f);
    return rc;
}

int cslam_pf_get_weights(cslam_pf_t* h, double* w) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_pf(h)) return rc;
    CSLAM_REQUIRE(w != nullptr, CSLAM_ERR_BAD_ARG, "null");
    CSLAM_CUDA(cudaMemcpyAsync(w, h->buf[h->cur].w, (size_t)h->np * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CSLAM_CUDA(cudaStreamSynchronize(h->stream));
    return CSLAM_OK;
}

static int pf_scratch(cslam_pf* h, size_t doubles) {  // accessor staging owned by the handle
    if (h->acc_cap >= doubles) return CSLAM_OK;
    CSLAM_CUDA(cudaStreamSynchronize(h->stream));
    cudaFree(h->acc_dev);
    h->acc_dev = nullptr;
    h->acc_cap = 0;
    const size_t want = std::max<size_t>(doubles, 4096);
    CSLAM_CUDA(cudaMalloc(&h->acc_dev, want * sizeof(double)));
    h->acc_cap = want;
    return CSLAM_OK;
}
static int get_aos(cslam_pf* h, const double* soa, int k, double* out) {
    const size_t cnt = (size_t)h->np * k;
    if (cnt == 0) return CSLAM_OK;
    if (int rc = pf_scratch(h, cnt)) return rc;
    count_launch();
    k_soa_to_aos<<<nblk(cnt, 256), 256, 0, h->stream>>>(soa, h->acc_dev, h->pp, h->np, k);
    CSLAM_CUDA(cudaGetLastError());
    CSLAM_CUDA(cudaMemcpyAsync(out, h->acc_dev, cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CSLAM_CUDA(cudaStreamSynchronize(h->stream));
    return CSLAM_OK;
}
static int set_aos(cslam_pf* h, double* soa, int k, const double* in) {
    const size_t cnt = (size_t)h->np * k;
    if (cnt == 0) return CSLAM_OK;
    if (int rc = pf_scratch(h, cnt)) return rc;
    CSLAM_CUDA(cudaMemcpyAsync(h->acc_dev, in, cnt * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    count_launch();
    k_aos_to_soa<<<nblk(cnt, 256), 256, 0, h->stream>>>(h->acc_dev, soa, h->pp, h->np, k);
    CSLAM_CUDA(cudaGetLastError());
    CSLAM_CUDA(cudaStreamSynchronize(h->stream));
    return CSLAM_OK;
}

int cslam_pf_get_poses(cslam_pf_t* h, double* X) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_pf(h)) return rc;
    CSLAM_REQUIRE(X != nullptr, CSLAM_ERR_BAD_ARG, "null");
    return get_aos(h, h->buf[h->cur].xv, 3, X);
}
int cslam_pf_get_pose_covs(cslam_pf_t* h, double* Pv) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_pf(h)) return rc;
    CSLAM_REQUIRE(Pv != nullptr, CSLAM_ERR_BAD_ARG, "null");
    return get_aos(h, h->buf[h->cur].pv, 9, Pv);
}
int cslam_pf_get_features(cslam_pf_t* h, int particle, double* XF, double* PF) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_pf(h)) return rc;
    CSLAM_REQUIRE(XF && PF && particle >= 0 && particle < h->np, CSLAM_ERR_BAD_ARG, "bad argument");
    if (h->nf == 0) return CSLAM_OK;
    if (int rc = pf_scratch(h, (size_t)6 * h->nf)) return rc;
    double* tmp = h->acc_dev;
    count_launch();
    k_get_features<<<nblk(h->nf, 128), 128, 0, h->stream>>>(h->buf[h->cur].xf, h->buf[h->cur].pf, h->pp, particle,
                                                            h->nf, tmp, tmp + 2 * h->nf);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(XF, tmp, (size_t)2 * h->nf * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(PF, tmp + 2 * h->nf, (size_t)4 * h->nf * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    CSLAM_CUDA(e);
    return CSLAM_OK;
}
// Slam::extractFeaturesFromParticles (slam.h:517-539): the feature estimates of EVERY particle, particle
// after particle — XF[p][f][2] (the reference concatenates the 2 x nf blocks column-wise in particle order) and,
// optionally, their packed 2x2 covariances PFp[p][f][3] = (xx, xy, yy).  Either pointer may be NULL.
int cslam_pf_get_features_all(cslam_pf_t* h, double* XF, double* PFp) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_pf(h)) return rc;
    if (h->nf == 0) return CSLAM_OK;
    if (XF)
        if (int rc = get_aos(h, h->buf[h->cur].xf, 2 * h->nf, XF)) return rc;
    if (PFp)
        if (int rc = get_aos(h, h->buf[h->cur].pf, 3 * h->nf, PFp)) return rc;
    return CSLAM_OK;
}
int cslam_pf_set_weights(cslam_pf_t* h, const double* w) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_pf(h)) return rc;
    CSLAM_REQUIRE(w != nullptr, CSLAM_ERR_BAD_ARG, "null");
    CSLAM_CUDA(cudaMemcpyAsync(h->buf[h->cur].w, w, (size_t)h->np * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CSLAM_CUDA(cudaStreamSynchronize(h->stream));
    return CSLAM_OK;
}
int cslam_pf_set_poses(cslam_pf_t* h, const double* X, const double* Pv) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_pf(h)) return rc;
    CSLAM_REQUIRE(X != nullptr, CSLAM_ERR_BAD_ARG, "null");
    if (int rc = set_aos(h, h->buf[h->cur].xv, 3, X)) return rc;
    if (Pv) return set_aos(h, h->buf[h->cur].pv, 9, Pv);
    return CSLAM_OK;
}

int cslam_pf_extract_state(cslam_pf_t* h, double X[3], int* index) {
    CSLAM_NVTX_RANGE();
    if (int rc = check_pf(h)) return rc;
    CSLAM_REQUIRE(X != nullptr, CSLAM_ERR_BAD_ARG, "null");
    const int nw = (h->np + 31) / 32;
    // candidates reuse the resampling scratch (wn: doubles, keep: ints), both idle between calls
    count_launch();
    k_argmin_w<<<nblk(h->np, 256), 256, 0, h->stream>>>(h->buf[h->cur].w, h->np, h->wn, h->keep);
    count_launch();
    k_argmin_final<<<1, 1024, 0, h->stream>>>(h->wn, h->keep, nw, h->buf[h->cur].xv, h->pp, h->d_small + 8);
    CSLAM_CUDA(cudaGetLastError());
    CSLAM_CUDA(cudaMemcpyAsync(h->pinned, h->d_small + 8, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CSLAM_CUDA(cudaStreamSynchronize(h->stream));
    const double* r = static_cast<const double*>(h->pinned);
    X[0] = r[0]; X[1] = r[1]; X[2] = r[2];
    if (index) *index = (int)r[3];
    return CSLAM_OK;
}

}  // extern "C"
