// oracle/oracle_capi.cpp — TEST INFRASTRUCTURE.  extern "C" face of slam_oracle.hpp (FP64
// instantiation) for ctypes: handle-based so that tests drive the oracle and the CUDA path
// with the same call sequence.  Also hosts the config-1 input-tape generator (the
// simulator half of test/main.cpp:24-201) and the CPU-baseline timing loops used by
// bench.py's cpu_baseline / --impl reference legs.  Never linked into libcslam.so.
#include <chrono>
#include <random>

#include "slam_oracle.hpp"


using namespace oracle;
// The same file is compiled twice: FP64 (orc_*, the parity oracle) and FP32 (orcf_*, which mimics
// the reference's own precision and is what tests/test_oracle_vs_ref.py pins against oracle/_ref).
#ifndef ORC_SCALAR
#define ORC_SCALAR double
#define ORC(name) orc_##name
#endif
typedef ORC_SCALAR T;

namespace {

struct EkfState {
    unsigned flags = 0;
    Vec<T> X = Vec<T>(3, 0.0);
    Mat<T> P = Mat<T>(3, 3);
};

struct PfState {
    unsigned flags = 0;
    std::vector<Particle<T>> ps;
};

Mat<T> mat2(const T* R) {  // 2x2, column-major == Eigen layout
    Mat<T> M(2, 2);
    M(0, 0) = R[0]; M(1, 0) = R[1]; M(0, 1) = R[2]; M(1, 1) = R[3];
    return M;
}
Mat<T> zmat(const T* Z, int m) {
    Mat<T> M(m > 0 ? 2 : 0, m);
    for (int i = 0; i < m; i++) {
        M(0, i) = Z[2 * i];
        M(1, i) = Z[2 * i + 1];
    }
    return M;
}

// test/main.cpp:24-54 landmark map and :67-76 waypoints (FP32 literals, as the reference stores them)
const float kLm1[30] = {1286.9623655913983384380117058754F,  -16.801075268817204301075268817204F,
                        2879.7043010752677218988537788391F,  4042.3387096774185920367017388344F,
                        2510.0806451612897944869473576546F,  -1871.6397849462364320061169564724F,
                        -2120.2956989247313686064444482327F, -3618.9516129032253957120701670647F,
                        -4210.3494623655915347626432776451F, -4317.8763440860211630933918058872F,
                        534.2741935483870967741935483871F,   -910.61827956989236554363742470741F,
                        -4290.9946236559135286370292305946F, 177.06919945726258447393774986267F,
                        1044.0976933514302800176665186882F,  506.78426051560745690949261188507F,
                        1813.4328358208986173849552869797F,  2656.0379918588914733845740556717F,
                        3242.1981004070585186127573251724F,  3999.3215739484458026709035038948F,
                        1532.5644504749034240376204252243F,  1117.3677069199529796605929732323F,
                        -152.64586160108228796161711215973F, -2008.8195386702818723279051482677F,
                        -3755.0881953867001357139088213444F, -3046.8113975576652592280879616737F,
                        -4902.9850746268630246049724519253F, 1654.6811397557721647899597883224F,
                        4194.7082767978317860979586839676F,  3278.8331071913198684342205524445F};
const float kLm2[30] = {203.82165605095541401273885350318F,  -1095.5414012738865494611673057079F,
                        -2942.6751592356704350095242261887F, -76.433121019108280254777070063694F,
                        3108.2802547770697856321930885315F,  4076.4331210191066929837688803673F,
                        191.08280254777070063694267515924F,  -3770.7006369426762830698862671852F,
                        -1235.6687898089185182470828294754F, 4089.1719745222908386494964361191F,
                        4789.8089171974515920737758278847F,  2420.3821656050940873683430254459F,
                        1286.6242038216551009099930524826F,  -164.38356164383561643835616438356F,
                        -1698.6301369863012951100245118141F, -1479.4520547945194266503676772118F,
                        -821.91780821917808219178082191781F, -630.13698630136986301369863013699F,
                        1041.0958904109589041095890410959F,  2054.7945205479445576202124357224F,
                        2219.1780821917818684596568346024F,  1369.863013698630136986301369863F,
                        1616.4383561643844586797058582306F,  2109.5890410958909342298284173012F,
                        1945.2054794520554423797875642776F,  1342.4657534246575342465753424658F,
                        1917.8082191780849825590848922729F,  -1616.4383561643826396903023123741F,
                        1150.6849315068493150684931506849F,  2000.0F};
const float kWp1[5] = {0.0F, 997.98387096774193548387096774194F, 4028.897849462364320061169564724F,
                       -1058.4677419354838709677419354839F, -4976.478494623655933537520468235F};
const float kWp2[5] = {0.0F, -2038.2165605095560749759897589684F, 1707.0063694267500977730378508568F,
                       1987.2611464968140353448688983917F, 1464.9681528662404161877930164337F};

}  // namespace

extern "C" {

// ------------------------------------------------------------------ EKF handle ----
void* ORC(ekf_create)(unsigned flags) {
    EkfState* s = new EkfState();
    s->flags = flags;
    return s;
}
void ORC(ekf_destroy)(void* h) { delete static_cast<EkfState*>(h); }
int ORC(ekf_n)(void* h) { return (int)static_cast<EkfState*>(h)->X.size(); }

void ORC(ekf_reset)(void* h, const T* X, int n, const T* P) {
    EkfState* s = static_cast<EkfState*>(h);
    s->X.assign(X, X + n);
    s->P = Mat<T>(n, n);
    if (P)
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) s->P(i, j) = P[(size_t)i * n + j];
}
void ORC(ekf_get_state)(void* h, T* X) {
    EkfState* s = static_cast<EkfState*>(h);
    std::copy(s->X.begin(), s->X.end(), X);
}
void ORC(ekf_get_cov)(void* h, T* P) {  // row-major n x n
    EkfState* s = static_cast<EkfState*>(h);
    const int n = s->P.r;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) P[(size_t)i * n + j] = s->P(i, j);
}
void ORC(ekf_predict)(void* h, double v, double swa, const T* Q, double wb, double dt) {
    EkfState* s = static_cast<EkfState*>(h);
    ekf_predict<T>(s->X, s->P, v, swa, mat2(Q), wb, dt, s->flags);
}
void ORC(ekf_observe_heading)(void* h, double phi, int use_heading, int dense) {
    EkfState* s = static_cast<EkfState*>(h);
    ekf_observe_heading<T>(s->X, s->P, phi, use_heading != 0, dense != 0);
}
int ORC(ekf_update)(void* h, const T* Z, const int* idf, int m, const T* R, int batch) {
    EkfState* s = static_cast<EkfState*>(h);
    std::vector<int> ids(idf, idf + m);
    return ekf_update<T>(s->X, s->P, zmat(Z, m), mat2(R), ids, batch != 0, s->flags);
}
void ORC(ekf_augment)(void* h, const T* Z, int m, const T* R) {
    EkfState* s = static_cast<EkfState*>(h);
    ekf_augment<T>(s->X, s->P, zmat(Z, m), mat2(R));
}
// returns #associated; idf_out (capacity m) = the reference's IDF vector, zn_count = ZN.cols()
int ORC(ekf_gate)(void* h, const T* Z, int m, const T* R, double gate1, double gate2, int dense,
                 int* jbest, uint8_t* is_new, T* nbest, T* outer, int* idf_out, int* zn_count) {
    EkfState* s = static_cast<EkfState*>(h);
    Association<T> a = ekf_data_associate<T>(s->X, s->P, zmat(Z, m), mat2(R), gate1, gate2, s->flags, dense != 0);
    for (int i = 0; i < m; i++) {
        if (jbest) jbest[i] = a.jbest[i];
        if (is_new) is_new[i] = a.is_new[i];
        if (nbest) nbest[i] = a.nbest[i];
        if (outer) outer[i] = a.outer[i];
    }
    if (idf_out) std::copy(a.idf.begin(), a.idf.end(), idf_out);
    if (zn_count) *zn_count = a.ZN.c;
    return (int)a.idf.size();
}
// EKF.cpp:146-233 on a caller-owned table; outputs split indices (into Z's columns)
void ORC(ekf_table)(void* h, const int* idz, int m, int* table, int table_len, int* zf_cols, int* idf, int* n_zf,
                   int* zn_cols, int* n_zn) {
    EkfState* s = static_cast<EkfState*>(h);
    std::vector<int> tab(table, table + table_len), ids(idz, idz + m);
    Mat<T> Z(m > 0 ? 2 : 0, m);
    // column bookkeeping: re-derive which observation went where (same rule as the function)
    int nf = 0, nn = 0;
    for (int i = 0; i < m; i++) {
        if (tab[ids[i] - 1] == 0) zn_cols[nn++] = i; else zf_cols[nf++] = i;
    }
    Association<T> a = ekf_data_associate_table<T>(s->X, Z, ids, tab);
    std::copy(a.idf.begin(), a.idf.end(), idf);
    *n_zf = nf;
    *n_zn = nn;
    std::copy(tab.begin(), tab.end(), table);
}

// ------------------------------------------------------------------ simulator tape ----
// The filter-independent half of test/main.cpp:93-200: true-pose propagation, noisy
// controls and observations.  noise_seed == 0 reproduces mSwitchControlNoise =
// mSwitchSensorNoise = false; otherwise draws come from std::mt19937_64(noise_seed)
// (the reference's clock-seeded draws are not reproducible, so they are inputs).
// controls[s] = (vn, swan, phi_true); observation steps have obs_ptr[s+1] > obs_ptr[s] or
// obs_flag[s] = 1 with zero visible landmarks.
// slam.h:575-582 getObservations for a world of N landmarks (LM 2 x N column-major); tags = 1..N.
// Returns the number of visible landmarks; the first max_out are written (Z interleaved range, bearing).
int ORC(get_observations)(const T* X3, const T* LM, int N, double rmax, int max_out, T* Z, int* tags_out) {
    Vec<T> X(3);
    for (int i = 0; i < 3; i++) X[i] = X3[i];
    Mat<T> L(2, N);
    std::vector<int> tags(N);
    for (int i = 0; i < N; i++) {
        L(0, i) = LM[2 * (size_t)i];
        L(1, i) = LM[2 * (size_t)i + 1];
        tags[i] = i + 1;
    }
    Mat<T> Zm;
    std::vector<int> vis;
    get_observations<T>(X, L, tags, (T)rmax, Zm, vis);
    const int m = (int)vis.size();
    for (int k = 0; k < m && k < max_out; k++) {
        Z[2 * k] = Zm(0, k);
        Z[2 * k + 1] = Zm(1, k);
        tags_out[k] = vis[k];
    }
    return m;
}

int ORC(sim_tape)(int max_steps, unsigned long long noise_seed, T* controls, int* obs_flag, int* obs_ptr,
                 T* Zout, int* tags_out, int max_obs, T* consts /* [8] out */) {
    const T V = 83.33F, maxSWA = (float)(kPi / 4.0F), rateSWA = (float)(70.0F * kPi / 180.0F), wb = 73.0F;
    const double dt = 0.01;
    const T sigmaV = 0.3F, sigmaSWA = (float)(1.0F * kPi / 180.0F);
    const T maxRange = 2000.0F;
    const double dtObserve = 5.058F * dt;
    const T sigmaR = 0.1F, sigmaB = (float)(1.0F * kPi / 180.0F);
    const T atWaypoint = 1.0F;
    Mat<T> LM(2, 30), WP(2, 5);
    for (int i = 0; i < 30; i++) { LM(0, i) = kLm1[i]; LM(1, i) = kLm2[i]; }
    for (int i = 0; i < 5; i++) { WP(0, i) = kWp1[i]; WP(1, i) = kWp2[i]; }
    std::vector<int> tags(30);
    for (int i = 0; i < 30; i++) tags[i] = i + 1;
    const T Q00 = sigmaV * sigmaV, Q11 = sigmaSWA * sigmaSWA, R00 = sigmaR * sigmaR, R11 = sigmaB * sigmaB;
    if (consts) {
        consts[0] = Q00; consts[1] = Q11; consts[2] = R00; consts[3] = R11;
        consts[4] = wb; consts[5] = dt; consts[6] = V; consts[7] = maxRange;
    }
    std::mt19937_64 rng(noise_seed);
    std::normal_distribution<double> nd(0.0, 1.0);
    Vec<T> XTrue(3, 0.0);
    int iwp = 1;
    T swa = 0;
    double dtsum = 0;
    int step = 0, nobs = 0;
    obs_ptr[0] = 0;
    while (iwp <= WP.c && iwp > 0 && step < max_steps) {
        compute_swa<T>(XTrue, WP, iwp, atWaypoint, swa, rateSWA, maxSWA, (T)dt);
        vehicle_model<T>(XTrue, V, swa, wb, (T)dt);
        T vn = V, swan = swa;
        if (noise_seed) {
            vn = vn + nd(rng) * std::sqrt(Q00);
            swan = swan + nd(rng) * std::sqrt(Q11);
        }
        controls[3 * step] = vn;
        controls[3 * step + 1] = swan;
        controls[3 * step + 2] = XTrue[2];
        obs_flag[step] = 0;
        dtsum = dtsum + dt;
        if (dtsum >= dtObserve) {
            dtsum = 0;
            obs_flag[step] = 1;
            Mat<T> Z;
            std::vector<int> vis;
            get_observations<T>(XTrue, LM, tags, maxRange, Z, vis);
            for (int k = 0; k < Z.c && nobs < max_obs; k++) {
                T zr = Z(0, k), zb = Z(1, k);
                if (noise_seed) {
                    zr = zr + nd(rng) * std::sqrt(R00);
                    zb = zb + nd(rng) * std::sqrt(R11);
                }
                Zout[2 * nobs] = zr;
                Zout[2 * nobs + 1] = zb;
                tags_out[nobs] = vis[k];
                nobs++;
            }
        }
        obs_ptr[step + 1] = nobs;
        step++;
    }
    return step;
}

// ------------------------------------------------------------------ PF handle ----
void* ORC(pf_create)(int num_particles, unsigned flags) {
    PfState* s = new PfState();
    s->flags = flags;
    s->ps = pf_initialize_particles<T>(num_particles);
    return s;
}
void ORC(pf_destroy)(void* h) { delete static_cast<PfState*>(h); }
int ORC(pf_num_features)(void* h) { return static_cast<PfState*>(h)->ps[0].XF.c; }
void ORC(pf_predict)(void* h, double v, double swa, const T* Q, double wb, double dt) {
    PfState* s = static_cast<PfState*>(h);
    Mat<T> Qm = mat2(Q);
    for (auto& p : s->ps) pf_predict<T>(p, v, swa, Qm, wb, dt);
}
void ORC(pf_observe_heading)(void* h, double phi, int use_heading) {
    PfState* s = static_cast<PfState*>(h);
    for (auto& p : s->ps) pf_observe_heading<T>(p, phi, use_heading != 0);
}
void ORC(pf_sample_proposal)(void* h, const T* Z, const int* idf, int m, const T* R, const T* xi) {
    PfState* s = static_cast<PfState*>(h);
    Mat<T> Zm = zmat(Z, m), Rm = mat2(R);
    std::vector<int> ids(idf, idf + m);
    for (size_t p = 0; p < s->ps.size(); p++) pf_sample_proposal<T>(s->ps[p], Zm, ids, Rm, xi + 3 * p, s->flags);
}
void ORC(pf_feature_update)(void* h, const T* Z, const int* idf, int m, const T* R) {
    PfState* s = static_cast<PfState*>(h);
    Mat<T> Zm = zmat(Z, m), Rm = mat2(R);
    std::vector<int> ids(idf, idf + m);
    for (auto& p : s->ps) pf_feature_update<T>(p, Zm, ids, Rm, s->flags);
}
int ORC(pf_resample)(void* h, const T* u, double num_effective, int resample_on, int* keep, T* neff) {
    PfState* s = static_cast<PfState*>(h);
    Vec<T> uv(u, u + s->ps.size());
    Stratified<T> st;
    const bool did = pf_resample_particles<T>(s->ps, uv, num_effective, resample_on != 0, s->flags, &st);
    if (keep) std::copy(st.keep.begin(), st.keep.end(), keep);
    if (neff) *neff = st.neff;
    return did ? 1 : 0;
}
// stratifiedResample alone (PF.cpp:546-577) on a weight vector
void ORC(stratified_resample)(const T* w, const T* u, int len, unsigned flags, int* keep, T* neff,
                             T* cumw) {
    Vec<T> W(w, w + len), U(u, u + len);
    Stratified<T> st = pf_stratified_resample<T>(W, U, flags);
    std::copy(st.keep.begin(), st.keep.end(), keep);
    *neff = st.neff;
    if (cumw) std::copy(W.begin(), W.end(), cumw);
}
void ORC(pf_add_features)(void* h, const T* Z, int m, const T* R) {
    PfState* s = static_cast<PfState*>(h);
    Mat<T> Zm = zmat(Z, m), Rm = mat2(R);
    for (auto& p : s->ps) pf_add_new_features<T>(p, Zm, Rm);
}
void ORC(pf_sample_pose)(void* h, const T* xi) {  // test/main.cpp:319-325
    PfState* s = static_cast<PfState*>(h);
    for (size_t p = 0; p < s->ps.size(); p++) {
        Particle<T>& q = s->ps[p];
        Mat<T> L = cholesky_decomposition(q.P);
        Vec<T> XS(3);
        for (int i = 0; i < 3; i++) {
            T acc = 0;
            for (int k = 0; k < 3; k++) acc += L(i, k) * xi[3 * p + k];
            XS[i] = acc + q.X[i];
        }
        q.X = XS;
        q.P = Mat<T>(3, 3);
    }
}
void ORC(pf_get_weights)(void* h, T* w) {
    PfState* s = static_cast<PfState*>(h);
    for (size_t p = 0; p < s->ps.size(); p++) w[p] = s->ps[p].w;
}
void ORC(pf_get_poses)(void* h, T* X, T* Pv) {
    PfState* s = static_cast<PfState*>(h);
    for (size_t p = 0; p < s->ps.size(); p++) {
        for (int i = 0; i < 3; i++) X[3 * p + i] = s->ps[p].X[i];
        if (Pv)
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) Pv[9 * p + 3 * i + j] = s->ps[p].P(i, j);
    }
}
void ORC(pf_get_features)(void* h, int particle, T* XF, T* PF) {
    PfState* s = static_cast<PfState*>(h);
    const Particle<T>& q = s->ps[particle];
    for (int f = 0; f < q.XF.c; f++) {
        XF[2 * f] = q.XF(0, f);
        XF[2 * f + 1] = q.XF(1, f);
        for (int i = 0; i < 2; i++)
            for (int j = 0; j < 2; j++) PF[4 * f + 2 * i + j] = q.PF[f](i, j);
    }
}
void ORC(pf_set_weights)(void* h, const T* w) {
    PfState* s = static_cast<PfState*>(h);
    for (size_t p = 0; p < s->ps.size(); p++) s->ps[p].w = w[p];
}
void ORC(pf_set_poses)(void* h, const T* X, const T* Pv) {
    PfState* s = static_cast<PfState*>(h);
    for (size_t p = 0; p < s->ps.size(); p++) {
        for (int i = 0; i < 3; i++) s->ps[p].X[i] = X[3 * p + i];
        if (Pv)
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) s->ps[p].P(i, j) = Pv[9 * p + 3 * i + j];
    }
}
int ORC(pf_extract_state)(void* h, T* X) {  // slam.h:493-511 (Q13: the MINIMUM-weight particle)
    PfState* s = static_cast<PfState*>(h);
    size_t pos = 0;
    for (size_t p = 1; p < s->ps.size(); p++)
        if (s->ps[p].w < s->ps[pos].w) pos = p;
    for (int i = 0; i < 3; i++) X[i] = s->ps[pos].X[i];
    return (int)pos;
}

// small standalone pieces for known-answer tests
double ORC(pi2pi)(double a) { return (double)pi2pi<T>((T)a); }
float ORC(pi2pi_f)(float a) { return pi2pi<float>(a); }
int ORC(cholesky)(const T* M, int n, T* L) {  // column-major in/out; returns used-fallback flag
    Mat<T> A(n, n);
    std::copy(M, M + (size_t)n * n, A.a.begin());
    bool fb = false;
    Mat<T> R = cholesky_decomposition(A, &fb);
    std::copy(R.a.begin(), R.a.end(), L);
    return fb ? 1 : 0;
}
void ORC(inverse)(const T* M, int n, T* out, T* det) {
    Mat<T> A(n, n);
    std::copy(M, M + (size_t)n * n, A.a.begin());
    Mat<T> I = inverse(A);
    std::copy(I.a.begin(), I.a.end(), out);
    if (det) *det = determinant(A);
}

}  // extern "C"
