// oracle/ref_capi.cpp — TEST INFRASTRUCTURE.  extern "C" face over the REFERENCE'S OWN classes
// (`EKF`, `PF` from /root/reference/slam, compiled unmodified where they lie, see oracle/Makefile
// target _ref) so that tests/test_oracle_vs_ref.py can pin oracle/slam_oracle.hpp against them.
// Arithmetic is FP32, exactly as the reference computes (Eigen::MatrixXf / float).
// Nothing here is copied from the reference: this file only CALLS its public methods.
#include <cstring>
#include <vector>

#include "EKF.h"
#include "PF.h"

namespace {
using Eigen::MatrixXf;
using Eigen::VectorXf;
using Eigen::VectorXi;

struct RefEkf {
    EKF* f;
    VectorXf X;
    MatrixXf P;
};
struct RefPf {
    PF* f;
    std::vector<Slam::Particle_t> ps;
};
MatrixXf m2(const float* R) {  // column-major 2x2
    MatrixXf M(2, 2);
    M(0, 0) = R[0]; M(1, 0) = R[1]; M(0, 1) = R[2]; M(1, 1) = R[3];
    return M;
}
MatrixXf zm(const float* Z, int m) {
    MatrixXf M(m > 0 ? 2 : 0, m);
    for (int i = 0; i < m; i++) { M(0, i) = Z[2 * i]; M(1, i) = Z[2 * i + 1]; }
    return M;
}
VectorXi iv(const int* a, int m) {
    VectorXi v = VectorXi::Zero(m);
    for (int i = 0; i < m; i++) v(i) = a[i];
    return v;
}
MatrixXf dummy_lm() { return MatrixXf::Zero(2, 30); }
MatrixXf dummy_wp() { return MatrixXf::Zero(2, 5); }
}  // namespace

extern "C" {

// ------------------------------------------------------------------ EKF ----
void* ref_ekf_create() {
    RefEkf* r = new RefEkf();
    r->f = new EKF(dummy_lm(), dummy_wp());
    r->X = VectorXf::Zero(3);
    r->P = MatrixXf::Zero(3, 3);
    return r;
}
void ref_ekf_destroy(void* h) {
    RefEkf* r = static_cast<RefEkf*>(h);
    delete r->f;
    delete r;
}
int ref_ekf_n(void* h) { return static_cast<RefEkf*>(h)->X.rows(); }
void ref_ekf_reset(void* h, const float* X, int n, const float* P) {
    RefEkf* r = static_cast<RefEkf*>(h);
    r->X = VectorXf::Zero(n);
    r->P = MatrixXf::Zero(n, n);
    for (int i = 0; i < n; i++) r->X(i) = X[i];
    if (P)
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) r->P(i, j) = P[(size_t)i * n + j];
}
void ref_ekf_get_state(void* h, float* X) {
    RefEkf* r = static_cast<RefEkf*>(h);
    for (int i = 0; i < r->X.rows(); i++) X[i] = r->X(i);
}
void ref_ekf_get_cov(void* h, float* P) {
    RefEkf* r = static_cast<RefEkf*>(h);
    const int n = r->P.rows();
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) P[(size_t)i * n + j] = r->P(i, j);
}
void ref_ekf_predict(void* h, double v, double swa, const float* Q, double wb, double dt) {
    RefEkf* r = static_cast<RefEkf*>(h);
    r->f->predict(r->X, r->P, (float)v, (float)swa, m2(Q), (float)wb, (float)dt);
}
void ref_ekf_observe_heading(void* h, double phi, int use_heading, int /*dense*/) {
    RefEkf* r = static_cast<RefEkf*>(h);
    r->f->observeHeading(r->X, r->P, (float)phi, use_heading != 0);
}
int ref_ekf_update(void* h, const float* Z, const int* idf, int m, const float* R, int batch) {
    RefEkf* r = static_cast<RefEkf*>(h);
    r->f->update(r->X, r->P, zm(Z, m), m2(R), iv(idf, m), batch != 0);
    return 0;
}
void ref_ekf_augment(void* h, const float* Z, int m, const float* R) {
    RefEkf* r = static_cast<RefEkf*>(h);
    r->f->augment(r->X, r->P, zm(Z, m), m2(R));
}
// dataAssociate (EKF.cpp:235-326): returns #associated, idf_out, ZN.cols(); nis/nd per (obs, landmark)
// are available through ref_ekf_compute_association.
int ref_ekf_data_associate(void* h, const float* Z, int m, const float* R, double gate1, double gate2,
                           int* idf_out, float* zf_out, int* zn_cols) {
    RefEkf* r = static_cast<RefEkf*>(h);
    auto a = r->f->dataAssociate(r->X, r->P, zm(Z, m), m2(R), (float)gate1, (float)gate2);
    for (int i = 0; i < a.idf.rows(); i++) idf_out[i] = a.idf(i);
    for (int i = 0; i < a.ZF.cols(); i++) { zf_out[2 * i] = a.ZF(0, i); zf_out[2 * i + 1] = a.ZF(1, i); }
    *zn_cols = a.ZN.cols();
    return a.idf.rows();
}
void ref_ekf_compute_association(void* h, const float* z, const float* R, int idf, float* nis, float* nd) {
    RefEkf* r = static_cast<RefEkf*>(h);
    auto ni = r->f->computeAssociation(r->X, r->P, zm(z, 1), m2(R), idf);
    *nis = ni.nis;
    *nd = ni.nd;
}
void ref_ekf_observe_model(void* h, int idf, float* z, float* H /* 2 x n row-major */) {
    RefEkf* r = static_cast<RefEkf*>(h);
    auto om = r->f->observeModel(r->X, idf);
    z[0] = om.Z(0, 0);
    z[1] = om.Z(1, 0);
    const int n = r->X.rows();
    for (int a = 0; a < 2; a++)
        for (int j = 0; j < n; j++) H[(size_t)a * n + j] = om.H(a, j);
}
// dataAssociateTable (EKF.cpp:146-233)
void ref_ekf_table(void* h, const float* Z, const int* idz, int m, int* table, int table_len, int* idf_out,
                   int* n_zf, int* n_zn) {
    RefEkf* r = static_cast<RefEkf*>(h);
    VectorXi tab = iv(table, table_len);
    auto a = r->f->dataAssociateTable(r->X, zm(Z, m), iv(idz, m), tab);
    for (int i = 0; i < a.idf.rows(); i++) idf_out[i] = a.idf(i);
    *n_zf = a.ZF.cols();
    *n_zn = a.ZN.cols();
    for (int i = 0; i < table_len; i++) table[i] = tab(i);
}

// ------------------------------------------------------------------ shared helpers ----
float ref_pi2pi(float a) {
    EKF f(dummy_lm(), dummy_wp());
    return f.pi2Pi(a);
}
int ref_cholesky(const float* M, int n, float* L) {  // column-major
    EKF f(dummy_lm(), dummy_wp());
    MatrixXf A(n, n);
    std::memcpy(A.data(), M, sizeof(float) * (size_t)n * n);
    MatrixXf R = f.choleskyDecomposition(A);
    std::memcpy(L, R.data(), sizeof(float) * (size_t)n * n);
    return 0;
}
// the three N(0,1) values slam.h:753-764 feeds every particle (engine re-seeded to 1 per call, Q7)
void ref_proposal_draws(float* xi) {
    PF f(dummy_lm(), dummy_wp());
    VectorXf mean = VectorXf::Zero(3);
    MatrixXf cov = MatrixXf::Identity(3, 3);
    MatrixXf s = f.multivariateNormalGaussianDistribution(mean, cov, 1);
    for (int i = 0; i < 3; i++) xi[i] = s(i, 0);
}
// simulator helpers (slam.h:279-332, :952-966, :575-582) for the config-1 input stream
void ref_compute_swa(const float* X, const float* WP /*2 x k col-major*/, int k, int* iwp, float minD, float* swa,
                     float rateSWA, float maxSWA, float dt) {
    EKF f(dummy_lm(), dummy_wp());
    VectorXf x = VectorXf::Zero(3);
    for (int i = 0; i < 3; i++) x(i) = X[i];
    MatrixXf wp(2, k);
    std::memcpy(wp.data(), WP, sizeof(float) * 2 * (size_t)k);
    f.computeSWA(x, wp, *iwp, minD, *swa, rateSWA, maxSWA, dt);
}
void ref_vehicle_model(float* X, float v, float swa, float wb, float dt) {
    EKF f(dummy_lm(), dummy_wp());
    VectorXf x = VectorXf::Zero(3);
    for (int i = 0; i < 3; i++) x(i) = X[i];
    f.vehicleModel(x, v, swa, wb, dt);
    for (int i = 0; i < 3; i++) X[i] = x(i);
}
int ref_get_observations(const float* X, const float* LM /*2 x k col-major*/, int k, float rmax, float* Z, int* tags) {
    EKF f(dummy_lm(), dummy_wp());
    VectorXf x = VectorXf::Zero(3);
    for (int i = 0; i < 3; i++) x(i) = X[i];
    MatrixXf lm(2, k);
    std::memcpy(lm.data(), LM, sizeof(float) * 2 * (size_t)k);
    VectorXi ids = VectorXi::Zero(k);
    for (int i = 0; i < k; i++) ids(i) = i + 1;
    auto o = f.getObservations(x, lm, ids, rmax);
    for (int i = 0; i < o.Z.cols(); i++) {
        Z[2 * i] = o.Z(0, i);
        Z[2 * i + 1] = o.Z(1, i);
        tags[i] = o.idf(i);
    }
    return o.Z.cols();
}

// ------------------------------------------------------------------ PF ----
void* ref_pf_create(int num_particles) {
    RefPf* r = new RefPf();
    r->f = new PF(dummy_lm(), dummy_wp());
    r->ps = r->f->initializeParticles(num_particles);
    return r;
}
void ref_pf_destroy(void* h) {
    RefPf* r = static_cast<RefPf*>(h);
    delete r->f;
    delete r;
}
int ref_pf_num_features(void* h) { return static_cast<RefPf*>(h)->ps[0].XF.cols(); }
void ref_pf_predict(void* h, double v, double swa, const float* Q, double wb, double dt) {
    RefPf* r = static_cast<RefPf*>(h);
    for (auto& p : r->ps) r->f->predict(p, (float)v, (float)swa, m2(Q), (float)wb, (float)dt);
}
void ref_pf_observe_heading(void* h, double phi, int use_heading) {
    RefPf* r = static_cast<RefPf*>(h);
    for (auto& p : r->ps) r->f->observeHeading(p, (float)phi, use_heading != 0);
}
void ref_pf_sample_proposal(void* h, const float* Z, const int* idf, int m, const float* R) {
    RefPf* r = static_cast<RefPf*>(h);
    for (auto& p : r->ps) r->f->sampleProposal(p, zm(Z, m), iv(idf, m), m2(R));
}
void ref_pf_feature_update(void* h, const float* Z, const int* idf, int m, const float* R) {
    RefPf* r = static_cast<RefPf*>(h);
    for (auto& p : r->ps) r->f->featureUpdate(p, zm(Z, m), iv(idf, m), m2(R));
}
void ref_pf_add_features(void* h, const float* Z, int m, const float* R) {
    RefPf* r = static_cast<RefPf*>(h);
    for (auto& p : r->ps) r->f->addOneNewFeature(p, zm(Z, m), m2(R));
}
void ref_pf_sample_pose(void* h) {  // test/main.cpp:319-325
    RefPf* r = static_cast<RefPf*>(h);
    for (auto& p : r->ps) {
        p.X = r->f->multivariateNormalGaussianDistribution(p.X, p.P, 1);
        p.P = MatrixXf::Zero(3, 3);
    }
}
// stratifiedResample (PF.cpp:546-577): select is clock-seeded random in the reference, so only neff,
// the cumulative weights and the CONSTANT-keep property (Q10) are comparable.
void ref_stratified_resample(const float* w, int len, float* keep, float* neff, float* cumw) {
    PF f(dummy_lm(), dummy_wp());
    MatrixXf W(1, len);
    for (int i = 0; i < len; i++) W(0, i) = w[i];
    auto st = f.stratifiedResample(W);
    for (int i = 0; i < len; i++) { keep[i] = st.keep(0, i); cumw[i] = W(0, i); }
    *neff = st.neff;
}
float ref_gauss_evaluate(const float* V, const float* S /*col-major DxD*/, int D) {
    PF f(dummy_lm(), dummy_wp());
    VectorXf v = VectorXf::Zero(D);
    for (int i = 0; i < D; i++) v(i) = V[i];
    MatrixXf s(D, D);
    std::memcpy(s.data(), S, sizeof(float) * (size_t)D * D);
    return f.gaussEvaluate(v, s, false);
}
void ref_pf_get_weights(void* h, float* w) {
    RefPf* r = static_cast<RefPf*>(h);
    for (size_t p = 0; p < r->ps.size(); p++) w[p] = r->ps[p].w;
}
void ref_pf_set_weights(void* h, const float* w) {
    RefPf* r = static_cast<RefPf*>(h);
    for (size_t p = 0; p < r->ps.size(); p++) r->ps[p].w = w[p];
}
void ref_pf_get_poses(void* h, float* X, float* Pv) {
    RefPf* r = static_cast<RefPf*>(h);
    for (size_t p = 0; p < r->ps.size(); p++) {
        for (int i = 0; i < 3; i++) X[3 * p + i] = r->ps[p].X(i);
        if (Pv)
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) Pv[9 * p + 3 * i + j] = r->ps[p].P(i, j);
    }
}
void ref_pf_set_poses(void* h, const float* X, const float* Pv) {
    RefPf* r = static_cast<RefPf*>(h);
    for (size_t p = 0; p < r->ps.size(); p++) {
        for (int i = 0; i < 3; i++) r->ps[p].X(i) = X[3 * p + i];
        if (Pv)
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) r->ps[p].P(i, j) = Pv[9 * p + 3 * i + j];
    }
}
void ref_pf_get_features(void* h, int particle, float* XF, float* PFo) {
    RefPf* r = static_cast<RefPf*>(h);
    const auto& q = r->ps[(size_t)particle];
    for (int f = 0; f < q.XF.cols(); f++) {
        XF[2 * f] = q.XF(0, f);
        XF[2 * f + 1] = q.XF(1, f);
        for (int i = 0; i < 2; i++)
            for (int j = 0; j < 2; j++) PFo[4 * f + 2 * i + j] = q.PF[(size_t)f](i, j);
    }
}
int ref_pf_extract_state(void* h, float* X) {
    RefPf* r = static_cast<RefPf*>(h);
    VectorXf x = r->f->extractStatesFromParticles(r->ps);
    for (int i = 0; i < 3; i++) X[i] = x(i);
    return 0;
}

}  // extern "C"
