// slam_gpu.hpp — C++ host adaptor (L1) over the C ABI (include/cslam.h): the reference's `Slam`
// filter interface (slam/include/slam.h) with the same method names, argument order and
// index conventions, for drivers written like test/main.cpp.
//
// Two layers:
//   * EkfGpuT<Vec, Mat> / PfGpuT<...> — templates over any dense vector / matrix type with the
//     Eigen-style surface  rows(), cols(), operator()(i[,j]), resize(...)  (Eigen::VectorXf /
//     MatrixXf, or the dependency-free cslam_host::DVec / DMat below).  The caller still OWNS
//     X and P exactly as in the reference (test/main.cpp:107-108): every call takes them by
//     reference and the adaptor writes the result back — X always, P when n <= p_writeback_max
//     (copying a 12.8 GB covariance per call is what the GPU path exists to avoid; call
//     fetchCovariance() explicitly for big maps).
//   * slam_gpu_eigen.hpp — `class EKFGpu : public Slam` overriding the reference's virtuals, usable
//     when the reference's own headers (and Eigen) are present: the only edit in test/main.cpp is
//     `new EKF(LM, WP)` -> `new EKFGpu(LM, WP)`.  The particle filter is exposed at
//     population level (PfGpuT below): the reference's per-particle virtuals are called in driver
//     loops over a host std::vector<Particle_t>, one PfGpuT call replaces one such loop.
//
// Error convention: the C ABI never throws; THIS adaptor does.  The reference swallows exceptions and
// prints (EKF.cpp:22-25 etc.), which is harmless there because a failed step leaves its host state as it
// was; here a failed device call (capacity exceeded, CUDA error) would leave the driver running on a filter
// that silently skipped a step, so every failure raises cslam_host::Error (std::runtime_error) with the
// library's message.  Numerically skipped updates (non-SPD S, slam.h:252-255) are NOT failures: they are
// counted and reported by skippedUpdates(), as in the reference they are silent.  There is no CPU fallback.
#pragma once
#include <cstdint>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "cslam.h"

namespace cslam_host {

// Minimal dense containers (column-major like Eigen) for drivers built without Eigen.
struct DVec {
    std::vector<double> a;
    DVec() = default;
    explicit DVec(int n) : a((size_t)n, 0.0) {}
    int rows() const { return (int)a.size(); }
    int cols() const { return 1; }
    int size() const { return (int)a.size(); }
    void resize(int n) { a.assign((size_t)n, 0.0); }
    double& operator()(int i) { return a[(size_t)i]; }
    double operator()(int i) const { return a[(size_t)i]; }
};
template <class T>
struct Mat {
    int r = 0, c = 0;
    std::vector<T> a;
    Mat() = default;
    Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, T(0)) {}
    int rows() const { return r; }
    int cols() const { return c; }
    int size() const { return r * c; }
    void resize(int r_, int c_) { r = r_; c = c_; a.assign((size_t)r_ * c_, T(0)); }
    T& operator()(int i, int j) { return a[(size_t)j * r + i]; }
    T operator()(int i, int j) const { return a[(size_t)j * r + i]; }
};
using DMat = Mat<double>;
struct IVec {
    std::vector<int> a;
    IVec() = default;
    explicit IVec(int n) : a((size_t)n, 0) {}
    int rows() const { return (int)a.size(); }
    int size() const { return (int)a.size(); }
    void resize(int n) { a.assign((size_t)n, 0); }
    int& operator()(int i) { return a[(size_t)i]; }
    int operator()(int i) const { return a[(size_t)i]; }
};

template <class M>
struct AssociationT {  // slam.h:438-443
    M ZF, ZN;
    std::vector<int> idf;
};

struct Error : std::runtime_error {
    int code;
    Error(int rc, const std::string& what) : std::runtime_error(what), code(rc) {}
};
inline void report(int rc, const char* where) {  // a failed device call is fatal for the step: raise
    if (rc != CSLAM_OK) throw Error(rc, std::string(where) + ": " + cslam_last_error());
}

template <class Vec, class MatT>
class EkfGpuT {
  public:
    // config fields with the reference's defaults (slam.h:65-103)
    double mVelocity = 83.33F, mMaxSWA = 3.14159265358979323846 / 4.0, mRateSWA = 70.0 * 3.14159265358979323846 / 180.0;
    double mWheelBase = 73.0F, mDtControls = 0.01;
    double mSigmaV = 0.3F, mSigmaSWA = 3.14159265358979323846 / 180.0;
    double mMaxRange = 2000.0F, mDtObserve = 5.058F * 0.01;
    double mSigmaR = 0.1F, mSigmaB = 3.14159265358979323846 / 180.0;
    double mGateReject = 50.0F, mGateAugment = 1000.0F;
    bool mSwitchHeadingKnown = true, mSwitchAssociationKnown = true, mSwitchBatchUpdate = true;
    std::vector<int> mTABLE;  // slam.h:105
    int p_writeback_max = 256;

    EkfGpuT(int num_map_landmarks, int capacity_landmarks, int device = 0, unsigned flags = CSLAM_FLAG_REF_LITERAL)
        : mTABLE((size_t)num_map_landmarks, 0), flags_(flags) {
        report(cslam_ekf_create(&h_, capacity_landmarks, device, flags), "EKFGpu");
    }
    ~EkfGpuT() { cslam_ekf_destroy(h_); }
    EkfGpuT(const EkfGpuT&) = delete;
    EkfGpuT& operator=(const EkfGpuT&) = delete;
    cslam_ekf_t* handle() { return h_; }

    // Slam::predict  slam.h:841-847 / EKF.cpp:406-455
    void predict(Vec& X, MatT& P, double v, double swa, const MatT& Q, double wb, double dt) {
        push(X, P);
        const double q[4] = {Q(0, 0), Q(1, 0), Q(0, 1), Q(1, 1)};
        report(cslam_ekf_predict(h_, v, swa, q, wb, dt), "predict");
        pull(X, P);
    }
    // Slam::observeHeading  slam.h:788 / EKF.cpp:328-352
    void observeHeading(Vec& X, MatT& P, double phi, bool useHeading = false) {
        push(X, P);
        report(cslam_ekf_observe_heading(h_, phi, useHeading ? 1 : 0), "observeHeading");
        pull(X, P);
    }
    // k control steps in one call: per step predict + observeHeading (test/main.cpp:140-168); the controls
    // and the measured heading come from the simulator / odometry and do not depend on the filter.
    // trace (nullable): 3*k doubles, X[0..2] after each step.
    void controlSteps(Vec& X, MatT& P, const std::vector<double>& v, const std::vector<double>& swa,
                      const std::vector<double>& phi, bool useHeading, const MatT& Q, double wb, double dt,
                      std::vector<double>* trace = nullptr) {
        push(X, P);
        const int k = (int)v.size();
        const double q[4] = {Q(0, 0), Q(1, 0), Q(0, 1), Q(1, 1)};
        if (trace) trace->assign((size_t)3 * k, 0.0);
        report(cslam_ekf_control_steps(h_, k, v.data(), swa.data(), phi.data(), useHeading ? 1 : 0, q, wb, dt,
                                       trace ? trace->data() : nullptr),
               "controlSteps");
        pull(X, P);
    }
    // Slam::update  slam.h:938-943 / EKF.cpp:481-496
    void update(Vec& X, MatT& P, const MatT& Z, const MatT& R, const std::vector<int>& idf, bool batch = false) {
        push(X, P);
        const int m = Z.cols();
        if (m > 0) {
            std::vector<double> z((size_t)2 * m);
            for (int i = 0; i < m; i++) { z[2 * i] = Z(0, i); z[2 * i + 1] = Z(1, i); }
            std::vector<int32_t> ids(idf.begin(), idf.end());
            const double r[4] = {R(0, 0), R(1, 0), R(0, 1), R(1, 1)};
            if (batch) {
                report(cslam_ekf_update(h_, z.data(), ids.data(), m, r, 1), "batchUpdate");
            } else {
                for (int b = 0; b < m; b += CSLAM_MAX_OBS) {
                    const int mc = (m - b < CSLAM_MAX_OBS) ? m - b : CSLAM_MAX_OBS;
                    report(cslam_ekf_update(h_, z.data() + 2 * b, ids.data() + b, mc, r, 0), "singleUpdate");
                }
            }
        }
        pull(X, P);
    }
    void singleUpdate(Vec& X, MatT& P, const MatT& Z, const MatT& R, const std::vector<int>& idf) {
        update(X, P, Z, R, idf, false);
    }
    void batchUpdate(Vec& X, MatT& P, const MatT& Z, const MatT& R, const std::vector<int>& idf) {
        update(X, P, Z, R, idf, true);
    }
    // Slam::augment  slam.h:190-191 / EKF.cpp:9-26 — resizes the caller's X and P as the reference does
    void augment(Vec& X, MatT& P, const MatT& Z, const MatT& R) {
        push(X, P);
        const int m = Z.cols();
        if (m > 0) {
            std::vector<double> z((size_t)2 * m);
            for (int i = 0; i < m; i++) { z[2 * i] = Z(0, i); z[2 * i + 1] = Z(1, i); }
            const double r[4] = {R(0, 0), R(1, 0), R(0, 1), R(1, 1)};
            for (int b = 0; b < m; b += CSLAM_MAX_OBS) {
                const int mc = (m - b < CSLAM_MAX_OBS) ? m - b : CSLAM_MAX_OBS;
                report(cslam_ekf_augment(h_, z.data() + 2 * b, mc, r), "augment");
            }
        }
        pull(X, P);
    }
    // update(X, P, ZF, R, idf, batch) immediately followed by augment(X, P, ZN, R) — test/main.cpp:188-189 — as one
    // call (cslam_ekf_observe_step): one single-CTA launch on small maps, bit-identical to the two calls.
    void observeStep(Vec& X, MatT& P, const MatT& ZF, const MatT& R, const std::vector<int>& idf, const MatT& ZN,
                     bool batch = true) {
        const int mf = ZF.cols(), mn = ZN.cols();
        if (mf > CSLAM_MAX_BATCH_OBS || mn > CSLAM_MAX_BATCH_OBS) {
            update(X, P, ZF, R, idf, batch);
            augment(X, P, ZN, R);
            return;
        }
        push(X, P);
        std::vector<double> zf((size_t)2 * mf), zn((size_t)2 * mn);
        for (int i = 0; i < mf; i++) { zf[2 * i] = ZF(0, i); zf[2 * i + 1] = ZF(1, i); }
        for (int i = 0; i < mn; i++) { zn[2 * i] = ZN(0, i); zn[2 * i + 1] = ZN(1, i); }
        std::vector<int32_t> ids(idf.begin(), idf.end());
        const double r[4] = {R(0, 0), R(1, 0), R(0, 1), R(1, 1)};
        report(cslam_ekf_observe_step(h_, mf ? zf.data() : nullptr, mf ? ids.data() : nullptr, mf,
                                      mn ? zn.data() : nullptr, mn, r, batch ? 1 : 0),
               "observeStep");
        pull(X, P);
    }
    // Slam::dataAssociate  slam.h:482-487 / EKF.cpp:235-326 (Q5: ZN empty unless CSLAM_FLAG_Q5_RETURN_ZN)
    AssociationT<MatT> dataAssociate(const Vec& X, const MatT& P, const MatT& Z, const MatT& R, double gate1,
                                    double gate2) {
        push(X, P);
        AssociationT<MatT> out;
        const int m = Z.cols();
        std::vector<double> z((size_t)2 * m);
        for (int i = 0; i < m; i++) { z[2 * i] = Z(0, i); z[2 * i + 1] = Z(1, i); }
        std::vector<int32_t> jbest((size_t)m, 0);
        std::vector<uint8_t> is_new((size_t)m, 0);
        const double r[4] = {R(0, 0), R(1, 0), R(0, 1), R(1, 1)};
        if (m > 0)
            report(cslam_ekf_gate(h_, z.data(), m, r, gate1, gate2, jbest.data(), is_new.data(), nullptr, nullptr),
                   "dataAssociate");
        int nzf = 0, nzn = 0;
        for (int i = 0; i < m; i++) { nzf += jbest[i] != 0; nzn += is_new[i] != 0; }
        out.ZF.resize(2, nzf);
        int k = 0;
        for (int i = 0; i < m; i++)
            if (jbest[i] != 0) { out.ZF(0, k) = Z(0, i); out.ZF(1, k) = Z(1, i); out.idf.push_back(jbest[i]); k++; }
        if (flags_ & CSLAM_FLAG_Q5_RETURN_ZN) {
            out.ZN.resize(2, nzn);
            k = 0;
            for (int i = 0; i < m; i++)
                if (is_new[i]) { out.ZN(0, k) = Z(0, i); out.ZN(1, k) = Z(1, i); k++; }
        } else {
            out.ZN.resize(0, 0);
        }
        return out;
    }
    // dataAssociate + update(batch = false) of the associated observations as ONE asynchronous
    // submission (test/main.cpp:193-195 without the host round trip; cslam_ekf_scan): the association
    // indices stay on the device.  Returns the same Association_t as dataAssociate (ZF, ZN, idf).
    AssociationT<MatT> scan(Vec& X, MatT& P, const MatT& Z, const MatT& R, double gate1, double gate2) {
        push(X, P);
        AssociationT<MatT> out;
        const int m = Z.cols();
        std::vector<double> z((size_t)2 * m);
        for (int i = 0; i < m; i++) { z[2 * i] = Z(0, i); z[2 * i + 1] = Z(1, i); }
        std::vector<int32_t> jbest((size_t)m, 0);
        std::vector<uint8_t> is_new((size_t)m, 0);
        const double r[4] = {R(0, 0), R(1, 0), R(0, 1), R(1, 1)};
        for (int b = 0; b < m; b += CSLAM_MAX_OBS) {
            const int mc = (m - b < CSLAM_MAX_OBS) ? m - b : CSLAM_MAX_OBS;
            report(cslam_ekf_scan(h_, z.data() + 2 * b, mc, r, gate1, gate2, jbest.data() + b, is_new.data() + b), "scan");
        }
        int nzf = 0, nzn = 0;
        for (int i = 0; i < m; i++) { nzf += jbest[i] != 0; nzn += is_new[i] != 0; }
        out.ZF.resize(2, nzf);
        int k = 0;
        for (int i = 0; i < m; i++)
            if (jbest[i] != 0) { out.ZF(0, k) = Z(0, i); out.ZF(1, k) = Z(1, i); out.idf.push_back(jbest[i]); k++; }
        if (flags_ & CSLAM_FLAG_Q5_RETURN_ZN) {
            out.ZN.resize(2, nzn);
            k = 0;
            for (int i = 0; i < m; i++)
                if (is_new[i]) { out.ZN(0, k) = Z(0, i); out.ZN(1, k) = Z(1, i); k++; }
        } else {
            out.ZN.resize(0, 0);
        }
        pull(X, P);
        return out;
    }
    // Slam::dataAssociateTable  slam.h:454-457 / EKF.cpp:146-233 (host bookkeeping)
    AssociationT<MatT> dataAssociateTable(const Vec& X, const MatT& Z, const std::vector<int>& idz,
                                         std::vector<int>& table) {
        AssociationT<MatT> out;
        std::vector<int> zf, zn, idn;
        for (size_t i = 0; i < idz.size(); i++) {
            const int id = idz[i];
            if (table[(size_t)id - 1] == 0) { zn.push_back((int)i); idn.push_back(id); }
            else { zf.push_back((int)i); out.idf.push_back(table[(size_t)id - 1]); }
        }
        out.ZF.resize(zf.empty() ? 0 : 2, (int)zf.size());
        for (size_t k = 0; k < zf.size(); k++) { out.ZF(0, (int)k) = Z(0, zf[k]); out.ZF(1, (int)k) = Z(1, zf[k]); }
        out.ZN.resize(zn.empty() ? 0 : 2, (int)zn.size());
        for (size_t k = 0; k < zn.size(); k++) { out.ZN(0, (int)k) = Z(0, zn[k]); out.ZN(1, (int)k) = Z(1, zn[k]); }
        const int nf = (X.rows() - 3) / 2;
        // the slots handed out here are only valid if the following augment() fits: check BEFORE touching the table
        if (nf + (int)idn.size() > cslam_ekf_capacity(h_))
            throw Error(CSLAM_ERR_CAPACITY, "dataAssociateTable: " + std::to_string(idn.size()) + " new landmarks do not fit "
                                            "the handle's capacity of " + std::to_string(cslam_ekf_capacity(h_)));
        for (size_t k = 0; k < idn.size(); k++) table[(size_t)idn[k] - 1] = nf + (int)k + 1;
        return out;
    }
    // The device copy of X, P is authoritative after the first call: the caller's X and P are OUTPUTS (X always,
    // P while n <= p_writeback_max).  A driver that edits its X / P in place (pose reset, covariance inflation)
    // must hand the new state over explicitly.
    void reseed(const Vec& X, const MatT& P) {
        seeded_ = false;
        push(X, P);
    }
    // true when P is no longer written back after each call (n > p_writeback_max): use fetchCovariance()
    bool covarianceWritebackSuppressed() const { return cslam_ekf_n(h_) > p_writeback_max; }
    // explicit covariance read-back for maps above p_writeback_max
    void fetchCovariance(MatT& P) {
        const int n = cslam_ekf_n(h_);
        std::vector<double> buf((size_t)n * n);
        report(cslam_ekf_get_cov_block(h_, 0, 0, n, n, buf.data()), "fetchCovariance");
        P.resize(n, n);
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) P(i, j) = buf[(size_t)i * n + j];
    }
    int skippedUpdates() {
        int s = 0;
        report(cslam_ekf_sync(h_, &s), "sync");
        return s;
    }

  private:
    // The caller owns X,P.  On the first call (or if the caller re-seeded them: size differs
    // from the device state) they are uploaded; afterwards the device copy is authoritative.
    void push(const Vec& X, const MatT& P) {
        const int n = X.rows();
        if (seeded_ && n == cslam_ekf_n(h_)) return;
        std::vector<double> x((size_t)n), p((size_t)n * n);
        for (int i = 0; i < n; i++) x[(size_t)i] = X(i);
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) p[(size_t)i * n + j] = P(i, j);
        report(cslam_ekf_reset(h_, x.data(), n, p.data()), "upload");
        seeded_ = true;
    }
    void pull(Vec& X, MatT& P) {
        const int n = cslam_ekf_n(h_);
        std::vector<double> x((size_t)n);
        report(cslam_ekf_get_state(h_, x.data(), n), "download");
        if (X.rows() != n) X.resize(n);
        for (int i = 0; i < n; i++) X(i) = x[(size_t)i];
        if (n <= p_writeback_max) {
            fetchCovariance(P);
        } else if (P.rows() != n) {
            P.resize(n, n);  // shape follows the reference's resize (EKF.cpp:69); values stay on the GPU
        }
    }
    cslam_ekf_t* h_ = nullptr;
    unsigned flags_ = 0;
    bool seeded_ = false;
};

using EkfGpu = EkfGpuT<DVec, DMat>;

// Particle filter (FastSLAM 2.0 style) — the PF half of the `Slam` interface at POPULATION level.
// The reference exposes per-particle virtuals and lets the driver loop over a host
// std::vector<Particle_t> (test/main.cpp:279-286, 305-309, 316-327); here the particles live on the
// GPU as a struct of arrays and each method below is one driver loop:
//
//     for (pl...) predict(particles[pl], vn, swan, QE, wb, dt);      ->  pf.predict(vn, swan, QE, wb, dt);
//     for (pl...) observeHeading(particles[pl], phi, known);         ->  pf.observeHeading(phi, known);
//     for (i...) { sampleProposal(particles[i], ZF, IDF, RE);        ->  pf.sampleProposal(ZF, IDF, RE, xi);
//                  featureUpdate(particles[i], ZF, IDF, RE); }       ->  pf.featureUpdate(ZF, IDF, RE);
//     resampleParticles(particles, mNumEffective, mSwitchResample);  ->  pf.resampleParticles(neff_min, u, on);
//     for (i...) { [X = multivariateNormal...(X, P, 1); P = 0;]      ->  pf.samplePose(xi);
//                  addOneNewFeature(particles[i], ZN, RE); }         ->  pf.addOneNewFeature(ZN, RE);
//     extractStatesFromParticles(particles)                          ->  pf.extractStatesFromParticles();
//
// Random draws are INPUTS: xi = 3 standard normals per particle ([p][3]),
// u = one deviate per resampling slot; the reference draws them inside slam.h:753-764 / PF.cpp:579-596.
template <class Vec, class MatT>
class PfGpuT {
  public:
    int mNumParticles = 100;
    double mNumEffective = 75.0;  // slam.h:92-93 (0.75 * 100, fixed at construction in the reference)
    std::vector<int> mTABLE;

    PfGpuT(int num_map_landmarks, int num_particles, int capacity_landmarks, int device = 0,
           unsigned flags = CSLAM_FLAG_REF_LITERAL)
        : mNumParticles(num_particles), mTABLE((size_t)num_map_landmarks, 0) {
        // Slam::initializeParticles(n)  slam.h:688 / PF.cpp:319-341
        report(cslam_pf_create(&h_, num_particles, capacity_landmarks, device, flags), "PFGpu");
    }
    ~PfGpuT() { cslam_pf_destroy(h_); }
    PfGpuT(const PfGpuT&) = delete;
    PfGpuT& operator=(const PfGpuT&) = delete;
    cslam_pf_t* handle() { return h_; }
    int numFeatures() const { return cslam_pf_num_features(h_); }  // particles[0].XF.cols(), main.cpp:299

    // Slam::predict(Particle_t&, ...) for every particle   slam.h:858-863 / PF.cpp:419-471
    void predict(double v, double swa, const MatT& Q, double wb, double dt) {
        const double q[4] = {Q(0, 0), Q(1, 0), Q(0, 1), Q(1, 1)};
        report(cslam_pf_predict(h_, v, swa, q, wb, dt), "predict");
    }
    // Slam::observeHeading(Particle_t&, phi, use)          slam.h:796 / PF.cpp:382-417
    void observeHeading(double phi, bool useHeading = false) {
        report(cslam_pf_observe_heading(h_, phi, useHeading ? 1 : 0), "observeHeading");
    }
    // k control steps (predict + observeHeading each) of every particle in one launch
    void controlSteps(const std::vector<double>& v, const std::vector<double>& swa, const std::vector<double>& phi,
                      bool useHeading, const MatT& Q, double wb, double dt) {
        const double q[4] = {Q(0, 0), Q(1, 0), Q(0, 1), Q(1, 1)};
        report(cslam_pf_control_steps(h_, (int)v.size(), v.data(), swa.data(), phi.data(), useHeading ? 1 : 0, q, wb,
                                      dt),
               "controlSteps");
    }
    // Slam::sampleProposal(Particle_t&, Z, idf, R)         slam.h:881-884 / PF.cpp:502-544
    void sampleProposal(const MatT& Z, const std::vector<int>& idf, const MatT& R, const std::vector<double>& xi) {
        const int m = Z.cols();
        if (m == 0) return;
        std::vector<double> z = flat(Z);
        std::vector<int32_t> ids(idf.begin(), idf.end());
        const double r[4] = {R(0, 0), R(1, 0), R(0, 1), R(1, 1)};
        report(cslam_pf_sample_proposal(h_, z.data(), ids.data(), m, r, xi.data(), 0), "sampleProposal");
    }
    // Slam::featureUpdate(Particle_t&, Z, idf, R)          slam.h:549-552 / PF.cpp:222-277
    void featureUpdate(const MatT& Z, const std::vector<int>& idf, const MatT& R) {
        const int m = Z.cols();
        if (m == 0) return;
        std::vector<double> z = flat(Z);
        std::vector<int32_t> ids(idf.begin(), idf.end());
        const double r[4] = {R(0, 0), R(1, 0), R(0, 1), R(1, 1)};
        report(cslam_pf_feature_update(h_, z.data(), ids.data(), m, r), "featureUpdate");
    }
    // Slam::resampleParticles(particles, numEffective, on)  slam.h:871-872 / PF.cpp:473-500; returns Keep
    // (0-based source index per slot, Q11), neff through the out-parameter.
    std::vector<int32_t> resampleParticles(double numEffective, const std::vector<double>& u, bool resampleStatus,
                                           double* neff = nullptr, bool* resampled = nullptr) {
        std::vector<int32_t> keep((size_t)mNumParticles, 0);
        double ne = 0.0;
        int did = 0;
        report(cslam_pf_resample(h_, u.data(), 0, numEffective, resampleStatus ? 1 : 0, keep.data(), &ne, &did),
               "resampleParticles");
        if (neff) *neff = ne;
        if (resampled) *resampled = did != 0;
        return keep;
    }
    // Slam::addOneNewFeature(Particle_t&, Z, R)            slam.h:134 / PF.cpp:9-60
    void addOneNewFeature(const MatT& Z, const MatT& R) {
        const int m = Z.cols();
        if (m == 0) return;
        std::vector<double> z = flat(Z);
        const double r[4] = {R(0, 0), R(1, 0), R(0, 1), R(1, 1)};
        for (int b = 0; b < m; b += CSLAM_MAX_OBS) {
            const int mc = (m - b < CSLAM_MAX_OBS) ? m - b : CSLAM_MAX_OBS;
            report(cslam_pf_add_features(h_, z.data() + 2 * b, mc, r), "addOneNewFeature");
        }
    }
    // particles[i].X = multivariateNormalGaussianDistribution(X, P, 1); P = 0     test/main.cpp:319-325
    void samplePose(const std::vector<double>& xi) { report(cslam_pf_sample_pose(h_, xi.data(), 0), "samplePose"); }
    // Slam::dataAssociateTable(Z, idz, table, nf)          slam.h:468-471 — EKF-style bookkeeping (Q8)
    AssociationT<MatT> dataAssociateTable(const MatT& Z, const std::vector<int>& idz, std::vector<int>& table, int nf) {
        AssociationT<MatT> out;
        std::vector<int> zf, zn, idn;
        for (size_t i = 0; i < idz.size(); i++) {
            const int id = idz[i];
            if (table[(size_t)id - 1] == 0) { zn.push_back((int)i); idn.push_back(id); }
            else { zf.push_back((int)i); out.idf.push_back(table[(size_t)id - 1]); }
        }
        out.ZF.resize(zf.empty() ? 0 : 2, (int)zf.size());
        for (size_t k = 0; k < zf.size(); k++) { out.ZF(0, (int)k) = Z(0, zf[k]); out.ZF(1, (int)k) = Z(1, zf[k]); }
        out.ZN.resize(zn.empty() ? 0 : 2, (int)zn.size());
        for (size_t k = 0; k < zn.size(); k++) { out.ZN(0, (int)k) = Z(0, zn[k]); out.ZN(1, (int)k) = Z(1, zn[k]); }
        for (size_t k = 0; k < idn.size(); k++) table[(size_t)idn[k] - 1] = nf + (int)k + 1;
        return out;
    }
    // Slam::extractStatesFromParticles  slam.h:493-511: pose of the MINIMUM-weight particle (Q13)
    Vec extractStatesFromParticles(int* index = nullptr) {
        double x[3] = {0, 0, 0};
        int idx = 0;
        report(cslam_pf_extract_state(h_, x, &idx), "extractStatesFromParticles");
        Vec out;
        out.resize(3);
        for (int i = 0; i < 3; i++) out(i) = x[i];
        if (index) *index = idx;
        return out;
    }
    std::vector<double> weights() {
        std::vector<double> w((size_t)mNumParticles);
        report(cslam_pf_get_weights(h_, w.data()), "weights");
        return w;
    }
    std::vector<double> poses() {  // [p][3]
        std::vector<double> x((size_t)3 * mNumParticles);
        report(cslam_pf_get_poses(h_, x.data()), "poses");
        return x;
    }
    int skippedUpdates() {
        int s = 0;
        report(cslam_pf_sync(h_, &s), "sync");
        return s;
    }

  private:
    static std::vector<double> flat(const MatT& Z) {
        const int m = Z.cols();
        std::vector<double> z((size_t)2 * m);
        for (int i = 0; i < m; i++) { z[2 * i] = Z(0, i); z[2 * i + 1] = Z(1, i); }
        return z;
    }
    cslam_pf_t* h_ = nullptr;
};

using PfGpu = PfGpuT<DVec, DMat>;

}  // namespace cslam_host
