// slam_gpu_eigen.hpp — the drop-in subclass for the reference's own driver.
//
//   #include "slam_gpu_eigen.hpp"
//   std::shared_ptr<Slam> ekfSlam(new EKFGpu(LM, WP));      // was: new EKF(LM, WP)   test/main.cpp:89
//
// `class EKFGpu : public Slam` overrides the reference's pure virtuals (slam/include/slam.h:134-943)
// with the exact signatures (Eigen::VectorXf / MatrixXf in and out, FP32 at the interface as in the
// reference, FP64 on the device) and forwards to the C ABI through EkfGpuT.  The particle-filter
// virtuals are stubbed the way the reference's own EKF.h stubs them (EKF.h:22-24,83-86,...).
// `class PFGpu : public Slam` (below) does the same for the particle-filter half: the reference's
// PER-PARTICLE virtuals keep their signatures, the population lives on the GPU.
// Requires the reference's headers (slam.h) and an Eigen-compatible <Eigen/Dense> on the include
// path, so it is compiled only where those exist (in this repo: against oracle/eigen_shim, see
// tests/test_dropin.py and INTEGRATION.md).
#pragma once
#include <functional>
#include <memory>
#include <random>
#include <stdexcept>

#include "slam.h"  // the reference's header
#include "slam_gpu.hpp"

class EKFGpu : public Slam {
  public:
    EKFGpu(const Eigen::MatrixXf& landMarks, const Eigen::MatrixXf& wayPoints, int capacity_landmarks = -1,
           int device = 0, unsigned flags = CSLAM_FLAG_REF_LITERAL)
        : Slam(landMarks, wayPoints),
          impl_((int)landMarks.cols(), capacity_landmarks < 0 ? (int)landMarks.cols() : capacity_landmarks, device,
                flags) {
        mTABLE = Eigen::VectorXi::Zero(getLandMarks().cols());  // EKF.cpp:6
    }
    ~EKFGpu() = default;

    // ---- EKF half of the interface ---------------------------------------------------------
    void predict(Eigen::VectorXf& X, Eigen::MatrixXf& P, const float& v, const float& swa, const Eigen::MatrixXf& Q,
                 const float& wb, const float& dt) override {
        impl_.predict(X, P, v, swa, Q, wb, dt);
    }
    void observeHeading(Eigen::VectorXf& X, Eigen::MatrixXf& P, const float& phi, bool useHeading = false) override {
        impl_.observeHeading(X, P, phi, useHeading);
    }
    void update(Eigen::VectorXf& X, Eigen::MatrixXf& P, const Eigen::MatrixXf& Z, const Eigen::MatrixXf& R,
                const Eigen::VectorXi& idf, bool batch = false) override {
        impl_.update(X, P, Z, R, ids(idf), batch);
    }
    void singleUpdate(Eigen::VectorXf& X, Eigen::MatrixXf& P, const Eigen::MatrixXf& Z, const Eigen::MatrixXf& R,
                      const Eigen::VectorXi& idf) override {
        impl_.update(X, P, Z, R, ids(idf), false);
    }
    void batchUpdate(Eigen::VectorXf& X, Eigen::MatrixXf& P, const Eigen::MatrixXf& Z, const Eigen::MatrixXf& R,
                     const Eigen::VectorXi& idf) override {
        impl_.update(X, P, Z, R, ids(idf), true);
    }
    void augment(Eigen::VectorXf& X, Eigen::MatrixXf& P, const Eigen::MatrixXf& Z, const Eigen::MatrixXf& R) override {
        impl_.augment(X, P, Z, R);
    }
    void addOneNewFeature(Eigen::VectorXf& X, Eigen::MatrixXf& P, const Eigen::MatrixXf& Z,
                          const Eigen::MatrixXf& R) override {
        impl_.augment(X, P, Z, R);
    }
    Association_t dataAssociate(const Eigen::VectorXf& X, const Eigen::MatrixXf& P, const Eigen::MatrixXf& Z,
                                const Eigen::MatrixXf& R, const float& gate1, const float& gate2) override {
        auto a = impl_.dataAssociate(X, P, Z, R, gate1, gate2);
        return {a.ZF, a.ZN, vec(a.idf)};
    }
    Association_t dataAssociateTable(const Eigen::VectorXf& X, const Eigen::MatrixXf& Z, const Eigen::VectorXi& idz,
                                     Eigen::VectorXi& table) override {
        std::vector<int> tab = ids(table);
        auto a = impl_.dataAssociateTable(X, Z, ids(idz), tab);
        for (int i = 0; i < (int)tab.size(); i++) table(i) = tab[(size_t)i];
        return {a.ZF, a.ZN, vec(a.idf)};
    }
    // helpers the reference exposes as virtuals but only uses internally: not on the GPU path
    NormalizedInnovation_t computeAssociation(const Eigen::VectorXf&, const Eigen::MatrixXf&, const Eigen::MatrixXf&,
                                              const Eigen::MatrixXf&, int) override { return {}; }
    ObserveModel_t observeModel(const Eigen::VectorXf&, int) override { return {}; }

    // ---- PF half: stubbed exactly like the reference's EKF.h does --------------------------
    void addOneNewFeature(Particle_t&, const Eigen::MatrixXf&, const Eigen::MatrixXf&) override {}
    Eigen::MatrixXf computeDelta(const Eigen::MatrixXf&, const Eigen::MatrixXf&) override { return {}; }
    Jacobians_t computeJacobians(const Particle_t&, const Eigen::VectorXi&, const Eigen::MatrixXf&) override { return {}; }
    Association_t dataAssociateTable(const Eigen::MatrixXf&, const Eigen::VectorXi&, Eigen::VectorXi&, int) override { return {}; }
    void featureUpdate(Particle_t&, const Eigen::MatrixXf&, const Eigen::VectorXi&, const Eigen::MatrixXf&) override {}
    float gaussEvaluate(const Eigen::VectorXf&, const Eigen::MatrixXf&, bool) override { return {}; }
    std::vector<Particle_t> initializeParticles(int) override { return {}; }
    float likelihood(const Particle_t&, const Eigen::MatrixXf&, const Eigen::VectorXi&, const Eigen::MatrixXf&) override { return {}; }
    Eigen::MatrixXf multivariateGauss(const Eigen::VectorXf&, const Eigen::MatrixXf&, int) override { return {}; }
    void observeHeading(Particle_t&, const float&, bool) override {}
    void predict(Particle_t&, const float&, const float&, const Eigen::MatrixXf&, const float&, const float&) override {}
    void resampleParticles(std::vector<Particle_t>&, int, bool) override {}
    void sampleProposal(Particle_t&, const Eigen::MatrixXf&, const Eigen::VectorXi&, const Eigen::MatrixXf&) override {}
    Stratified_t stratifiedResample(Eigen::MatrixXf&) override { return {}; }
    Eigen::MatrixXf stratifiedRandom(int) override { return {}; }

    cslam_host::EkfGpuT<Eigen::VectorXf, Eigen::MatrixXf>& impl() { return impl_; }

  private:
    static std::vector<int> ids(const Eigen::VectorXi& v) {
        std::vector<int> out((size_t)v.size());
        for (int i = 0; i < (int)v.size(); i++) out[(size_t)i] = v(i);
        return out;
    }
    static Eigen::VectorXi vec(const std::vector<int>& v) {
        Eigen::VectorXi out = Eigen::VectorXi::Zero((int)v.size());
        for (int i = 0; i < (int)v.size(); i++) out(i) = v[(size_t)i];
        return out;
    }
    cslam_host::EkfGpuT<Eigen::VectorXf, Eigen::MatrixXf> impl_;
};


// ---------------------------------------------------------------------------------------------------
//   std::shared_ptr<Slam> pfSlam(new PFGpu(LM, WP));        // was: new PF(LM, WP)   test/main.cpp:204
//
// The reference drives its particle filter through PER-PARTICLE virtuals in loops over a host
// std::vector<Particle_t> (test/main.cpp:279-286, 305-309, 316-327).  PFGpu keeps those signatures
// (slam.h:134, 549-552, 688, 796, 858-863, 871-872, 881-884) and runs the population on the GPU:
//   * initializeParticles(n) creates the device population and returns the host vector the driver owns;
//   * a per-particle call records its Particle_t* ; the n-th call of a loop (every particle visited once, as
//     the driver's loops do, all with the same arguments) runs ONE population kernel and writes every
//     recorded particle back (w, X, P always; XF / PF while particles x features <= feature_writeback_max,
//     beyond that only their shapes), so host reads such as particles[0].XF.cols() (main.cpp:299) and the
//     inline Slam::extractStatesFromParticles (slam.h:493-511) keep working unmodified;
//   * host-side edits of particles[i].X / .P between loops (main.cpp:319-324 samples the pose on the host)
//     are detected against the last write-back and uploaded before the next population kernel;
//   * resampleParticles is already population-level in the reference and maps 1:1.
// Random draws default to what the reference does — sampleProposal: the SAME three
// normals for every particle (slam.h:758-763 re-seeds its generator to 1 on every call; they are taken from
// the inherited multivariateNormalGaussianDistribution itself), resampling: clock-seeded normals
// (slam.h:587-594) — and become a reproducible mt19937_64 stream after seed(s).
// Failures of the device path THROW (std::runtime_error): a driver must not carry on with a filter that
// silently skipped a step.
class PFGpu : public Slam {
  public:
    size_t feature_writeback_max = (size_t)1 << 22;  // particles x features above which XF / PF values stay on the GPU

    PFGpu(const Eigen::MatrixXf& landMarks, const Eigen::MatrixXf& wayPoints, int capacity_landmarks = -1, int device = 0,
          unsigned flags = CSLAM_FLAG_REF_LITERAL)
        : Slam(landMarks, wayPoints),
          cap_(capacity_landmarks < 0 ? (int)landMarks.cols() : capacity_landmarks),
          device_(device),
          flags_(flags) {}
    ~PFGpu() { cslam_pf_destroy(h_); }
    PFGpu(const PFGpu&) = delete;
    PFGpu& operator=(const PFGpu&) = delete;
    cslam_pf_t* handle() { return h_; }
    void seed(unsigned long long s) { rng_.seed(s); seeded_ = true; }

    // ---- PF half of the interface ------------------------------------------------------------
    std::vector<Particle_t> initializeParticles(int numParticles) override {  // slam.h:688 / PF.cpp:319-341
        cslam_pf_destroy(h_);
        h_ = nullptr;
        ok(cslam_pf_create(&h_, numParticles, cap_, device_, flags_), "initializeParticles");
        n_ = numParticles;
        std::vector<Particle_t> particles;
        for (int i = 0; i < n_; i++) {
            Particle_t p = {};
            p.w = 1.0F / (float)n_;
            p.X = Eigen::VectorXf::Zero(3);
            p.P = Eigen::MatrixXf::Zero(3, 3);
            p.XF.resize(0, 0);
            particles.push_back(p);
        }
        mirror_fetch(false);
        for (Loop* l : {&l_predict_, &l_heading_, &l_sample_, &l_feature_, &l_add_}) l->seen.clear();
        return particles;
    }
    void predict(Particle_t& particle, const float& v, const float& swa, const Eigen::MatrixXf& Q, const float& wb,
                 const float& dt) override {  // slam.h:858-863 / PF.cpp:419-471
        if (!visit(l_predict_, particle)) return;
        const double q[4] = {Q(0, 0), Q(1, 0), Q(0, 1), Q(1, 1)};
        upload_if_edited(l_predict_);
        ok(cslam_pf_predict(h_, v, swa, q, wb, dt), "predict");
        finish(l_predict_, false);
    }
    void observeHeading(Particle_t& particle, const float& phi, bool useHeading = false) override {  // slam.h:796
        if (!visit(l_heading_, particle)) return;
        upload_if_edited(l_heading_);
        ok(cslam_pf_observe_heading(h_, phi, useHeading ? 1 : 0), "observeHeading");
        finish(l_heading_, false);
    }
    void sampleProposal(Particle_t& particle, const Eigen::MatrixXf& Z, const Eigen::VectorXi& idf,
                        const Eigen::MatrixXf& R) override {  // slam.h:881-884 / PF.cpp:502-544
        if (!visit(l_sample_, particle)) return;
        upload_if_edited(l_sample_);
        const int m = (int)Z.cols();
        if (m > 0) {
            std::vector<double> z = flat(Z);
            std::vector<int32_t> ids = ids32(idf);
            const double r[4] = {R(0, 0), R(1, 0), R(0, 1), R(1, 1)};
            std::vector<double> xi((size_t)3 * n_);
            draw_xi(xi);
            ok(cslam_pf_sample_proposal(h_, z.data(), ids.data(), m, r, xi.data(), 0), "sampleProposal");
        }
        finish(l_sample_, false);
    }
    void featureUpdate(Particle_t& particle, const Eigen::MatrixXf& Z, const Eigen::VectorXi& idf,
                       const Eigen::MatrixXf& R) override {  // slam.h:549-552 / PF.cpp:222-277
        if (!visit(l_feature_, particle)) return;
        upload_if_edited(l_feature_);
        const int m = (int)Z.cols();
        if (m > 0) {
            std::vector<double> z = flat(Z);
            std::vector<int32_t> ids = ids32(idf);
            const double r[4] = {R(0, 0), R(1, 0), R(0, 1), R(1, 1)};
            ok(cslam_pf_feature_update(h_, z.data(), ids.data(), m, r), "featureUpdate");
        }
        finish(l_feature_, true);
    }
    void addOneNewFeature(Particle_t& particle, const Eigen::MatrixXf& Z, const Eigen::MatrixXf& R) override {
        // slam.h:134 / PF.cpp:9-60
        if (!visit(l_add_, particle)) return;
        upload_if_edited(l_add_);
        const int m = (int)Z.cols();
        if (m > 0) {
            if (cslam_pf_num_features(h_) + m > cap_)
                throw std::runtime_error("PFGpu::addOneNewFeature: landmark capacity exceeded");
            std::vector<double> z = flat(Z);
            const double r[4] = {R(0, 0), R(1, 0), R(0, 1), R(1, 1)};
            for (int b = 0; b < m; b += CSLAM_MAX_OBS) {
                const int mc = (m - b < CSLAM_MAX_OBS) ? m - b : CSLAM_MAX_OBS;
                ok(cslam_pf_add_features(h_, z.data() + 2 * b, mc, r), "addOneNewFeature");
            }
        }
        finish(l_add_, true);
    }
    void resampleParticles(std::vector<Particle_t>& particles, int numEffective, bool resampleStatus = false) override {
        // slam.h:871-872 / PF.cpp:473-500 -> stratifiedResample :546-577 (+ :579-596 for the deviates)
        if ((int)particles.size() != n_) throw std::runtime_error("PFGpu::resampleParticles: particle count changed");
        Loop all;
        for (auto& p : particles) all.seen.push_back(&p);
        upload_if_edited(all);
        std::vector<double> u((size_t)n_);
        for (auto& x : u) x = seeded_ ? normal_(rng_) : (double)generateRandomNumber<float>();
        int did = 0;
        double ne = 0.0;
        ok(cslam_pf_resample(h_, u.data(), 0, (double)numEffective, resampleStatus ? 1 : 0, nullptr, &ne, &did),
           "resampleParticles");
        last_neff_ = ne;
        finish(all, did != 0);
    }
    Association_t dataAssociateTable(const Eigen::MatrixXf& Z, const Eigen::VectorXi& idz, Eigen::VectorXi& table,
                                     int nf) override {
        // slam.h:468-471.  The reference's PF.cpp:204-213 walks the wrong id list and throws; this is
        // the bookkeeping it was transliterated from (= EKF.cpp:212-226), as PfGpuT does.
        std::vector<int> zf, zn, idn, idf;
        for (int i = 0; i < (int)idz.size(); i++) {
            const int id = idz(i);
            if (table(id - 1) == 0) { zn.push_back(i); idn.push_back(id); }
            else { zf.push_back(i); idf.push_back(table(id - 1)); }
        }
        Association_t out;
        out.ZF = Eigen::MatrixXf::Zero(zf.empty() ? 0 : 2, (int)zf.size());
        for (int k = 0; k < (int)zf.size(); k++) { out.ZF(0, k) = Z(0, zf[(size_t)k]); out.ZF(1, k) = Z(1, zf[(size_t)k]); }
        out.ZN = Eigen::MatrixXf::Zero(zn.empty() ? 0 : 2, (int)zn.size());
        for (int k = 0; k < (int)zn.size(); k++) { out.ZN(0, k) = Z(0, zn[(size_t)k]); out.ZN(1, k) = Z(1, zn[(size_t)k]); }
        out.idf = Eigen::VectorXi::Zero((int)idf.size());
        for (int k = 0; k < (int)idf.size(); k++) out.idf(k) = idf[(size_t)k];
        for (int k = 0; k < (int)idn.size(); k++) table(idn[(size_t)k] - 1) = nf + k + 1;
        return out;
    }
    double lastNeff() const { return last_neff_; }

    // helpers the reference exposes as virtuals but only uses inside PF.cpp: not on the GPU path
    Eigen::MatrixXf computeDelta(const Eigen::MatrixXf&, const Eigen::MatrixXf&) override { return {}; }
    Jacobians_t computeJacobians(const Particle_t&, const Eigen::VectorXi&, const Eigen::MatrixXf&) override { return {}; }
    float gaussEvaluate(const Eigen::VectorXf&, const Eigen::MatrixXf&, bool) override { return {}; }
    float likelihood(const Particle_t&, const Eigen::MatrixXf&, const Eigen::VectorXi&, const Eigen::MatrixXf&) override { return {}; }
    Eigen::MatrixXf multivariateGauss(const Eigen::VectorXf&, const Eigen::MatrixXf&, int) override { return {}; }
    Stratified_t stratifiedResample(Eigen::MatrixXf&) override { return {}; }
    Eigen::MatrixXf stratifiedRandom(int) override { return {}; }

    // ---- EKF half: stubbed exactly like the reference's PF.h does ------------------------------
    void augment(Eigen::VectorXf&, Eigen::MatrixXf&, const Eigen::MatrixXf&, const Eigen::MatrixXf&) override {}
    void addOneNewFeature(Eigen::VectorXf&, Eigen::MatrixXf&, const Eigen::MatrixXf&, const Eigen::MatrixXf&) override {}
    void batchUpdate(Eigen::VectorXf&, Eigen::MatrixXf&, const Eigen::MatrixXf&, const Eigen::MatrixXf&,
                     const Eigen::VectorXi&) override {}
    NormalizedInnovation_t computeAssociation(const Eigen::VectorXf&, const Eigen::MatrixXf&, const Eigen::MatrixXf&,
                                              const Eigen::MatrixXf&, int) override { return {}; }
    Association_t dataAssociateTable(const Eigen::VectorXf&, const Eigen::MatrixXf&, const Eigen::VectorXi&,
                                     Eigen::VectorXi&) override { return {}; }
    Association_t dataAssociate(const Eigen::VectorXf&, const Eigen::MatrixXf&, const Eigen::MatrixXf&,
                                const Eigen::MatrixXf&, const float&, const float&) override { return {}; }
    void observeHeading(Eigen::VectorXf&, Eigen::MatrixXf&, const float&, bool) override {}
    ObserveModel_t observeModel(const Eigen::VectorXf&, int) override { return {}; }
    void predict(Eigen::VectorXf&, Eigen::MatrixXf&, const float&, const float&, const Eigen::MatrixXf&, const float&,
                 const float&) override {}
    void singleUpdate(Eigen::VectorXf&, Eigen::MatrixXf&, const Eigen::MatrixXf&, const Eigen::MatrixXf&,
                      const Eigen::VectorXi&) override {}
    void update(Eigen::VectorXf&, Eigen::MatrixXf&, const Eigen::MatrixXf&, const Eigen::MatrixXf&, const Eigen::VectorXi&,
                bool) override {}

  private:
    struct Loop {
        std::vector<Particle_t*> seen;  // particles visited by the running driver loop, in call order
    };
    static void ok(int rc, const char* where) {
        if (rc != CSLAM_OK) throw std::runtime_error(std::string("PFGpu::") + where + ": " + cslam_last_error());
    }
    // records the call; true on the n-th call of the loop (= run the population kernel now)
    bool visit(Loop& l, Particle_t& p) {
        if (!h_) throw std::runtime_error("PFGpu: initializeParticles() has not been called");
        l.seen.push_back(&p);
        return (int)l.seen.size() == n_;
    }
    // main.cpp:319-324 edits particles[i].X / .P on the host between loops: upload what differs from the last
    // write-back (slot k of a loop is particle k of the population: the driver walks its vector in order)
    void upload_if_edited(const Loop& l) {
        bool edited = false;
        for (int k = 0; k < n_ && !edited; k++) {
            const Particle_t& p = *l.seen[(size_t)k];
            for (int a = 0; a < 3 && !edited; a++) {
                if (p.X.rows() != 3 || p.X(a) != (float)mx_[(size_t)3 * k + a]) edited = true;
                for (int b = 0; b < 3 && !edited; b++)
                    if (p.P.rows() != 3 || p.P.cols() != 3 || p.P(a, b) != (float)mp_[(size_t)9 * k + 3 * a + b]) edited = true;
            }
        }
        if (!edited) return;
        // the device keeps FP64 state: only particles whose FP32 host image changed are replaced by the host values
        for (int k = 0; k < n_; k++) {
            const Particle_t& p = *l.seen[(size_t)k];
            bool ch = false;
            for (int a = 0; a < 3; a++) {
                ch = ch || p.X(a) != (float)mx_[(size_t)3 * k + a];
                for (int b = 0; b < 3; b++) ch = ch || p.P(a, b) != (float)mp_[(size_t)9 * k + 3 * a + b];
            }
            if (!ch) continue;
            for (int a = 0; a < 3; a++) {
                mx_[(size_t)3 * k + a] = p.X(a);
                for (int b = 0; b < 3; b++) mp_[(size_t)9 * k + 3 * a + b] = p.P(a, b);
            }
        }
        ok(cslam_pf_set_poses(h_, mx_.data(), mp_.data()), "upload of host-edited poses");
    }
    void mirror_fetch(bool features) {
        mw_.resize((size_t)n_);
        mx_.resize((size_t)3 * n_);
        mp_.resize((size_t)9 * n_);
        ok(cslam_pf_get_weights(h_, mw_.data()), "weights");
        ok(cslam_pf_get_poses(h_, mx_.data()), "poses");
        ok(cslam_pf_get_pose_covs(h_, mp_.data()), "pose covariances");
        nf_ = cslam_pf_num_features(h_);
        feat_values_ = (size_t)n_ * (size_t)nf_ <= feature_writeback_max;
        if (features && feat_values_ && nf_ > 0) {
            mxf_.resize((size_t)n_ * nf_ * 2);
            mpf_.resize((size_t)n_ * nf_ * 3);
            ok(cslam_pf_get_features_all(h_, mxf_.data(), mpf_.data()), "features");
            feat_fresh_ = true;
        } else if (features) {
            feat_fresh_ = false;
        }
    }
    // population kernel done: fetch the small per-particle state once and write every visited particle back
    void finish(Loop& l, bool features_changed) {
        mirror_fetch(features_changed);
        for (int k = 0; k < n_; k++) {
            Particle_t& p = *l.seen[(size_t)k];
            p.w = (float)mw_[(size_t)k];
            if (p.X.rows() != 3) p.X = Eigen::VectorXf::Zero(3);
            if (p.P.rows() != 3 || p.P.cols() != 3) p.P = Eigen::MatrixXf::Zero(3, 3);
            for (int a = 0; a < 3; a++) {
                p.X(a) = (float)mx_[(size_t)3 * k + a];
                for (int b = 0; b < 3; b++) p.P(a, b) = (float)mp_[(size_t)9 * k + 3 * a + b];
            }
            if (!features_changed) continue;
            if (p.XF.cols() != nf_) p.XF = Eigen::MatrixXf::Zero(nf_ > 0 ? 2 : 0, nf_);
            if ((int)p.PF.size() != nf_) p.PF.assign((size_t)nf_, Eigen::MatrixXf::Zero(2, 2));
            if (!(feat_values_ && feat_fresh_)) continue;
            for (int f = 0; f < nf_; f++) {
                const size_t o = ((size_t)k * nf_ + f);
                p.XF(0, f) = (float)mxf_[2 * o];
                p.XF(1, f) = (float)mxf_[2 * o + 1];
                Eigen::MatrixXf& c = p.PF[(size_t)f];
                c(0, 0) = (float)mpf_[3 * o];
                c(0, 1) = c(1, 0) = (float)mpf_[3 * o + 1];
                c(1, 1) = (float)mpf_[3 * o + 2];
            }
        }
        l.seen.clear();
    }
    void draw_xi(std::vector<double>& xi) {
        if (seeded_) {
            for (auto& x : xi) x = normal_(rng_);
            return;
        }
        // the reference's own draw: chol(I) * xi + 0 with its generator re-seeded to 1 (slam.h:753-764, Q7)
        const Eigen::MatrixXf s =
            multivariateNormalGaussianDistribution(Eigen::VectorXf::Zero(3), Eigen::MatrixXf::Identity(3, 3), 1);
        for (int k = 0; k < n_; k++)
            for (int a = 0; a < 3; a++) xi[(size_t)3 * k + a] = s(a, 0);
    }
    static std::vector<double> flat(const Eigen::MatrixXf& Z) {
        const int m = (int)Z.cols();
        std::vector<double> z((size_t)2 * m);
        for (int i = 0; i < m; i++) { z[2 * i] = Z(0, i); z[2 * i + 1] = Z(1, i); }
        return z;
    }
    static std::vector<int32_t> ids32(const Eigen::VectorXi& v) {
        std::vector<int32_t> out((size_t)v.size());
        for (int i = 0; i < (int)v.size(); i++) out[(size_t)i] = v(i);
        return out;
    }
    cslam_pf_t* h_ = nullptr;
    int n_ = 0, nf_ = 0, cap_ = 0, device_ = 0;
    unsigned flags_ = 0;
    Loop l_predict_, l_heading_, l_sample_, l_feature_, l_add_;
    std::vector<double> mw_, mx_, mp_, mxf_, mpf_;  // host image of the last write-back (FP64 as fetched)
    bool feat_values_ = true, feat_fresh_ = false;
    bool seeded_ = false;
    std::mt19937_64 rng_{1};
    std::normal_distribution<double> normal_{0.0, 1.0};
    double last_neff_ = 0.0;
};
