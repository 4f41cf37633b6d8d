// slam_gpu_eigen.hpp — the drop-in subclass for the reference's own driver.
//
//   #include "slam_gpu_eigen.hpp"
//   std::shared_ptr<Slam> ekfSlam(new EKFGpu(LM, WP));      // was: new EKF(LM, WP)   test/main.cpp:89
//
// `class EKFGpu : public Slam` overrides the reference's pure virtuals (slam/include/slam.h:134-943)
// with the exact signatures (Eigen::VectorXf / MatrixXf in and out, FP32 at the interface as in the
// reference, FP64 on the device) and forwards to the C ABI through EkfGpuT.  The particle-filter
// virtuals are stubbed the way the reference's own EKF.h stubs them (EKF.h:22-24,83-86,...).
// Requires the reference's headers (slam.h) and an Eigen-compatible <Eigen/Dense> on the include
// path, so it is compiled only where those exist (in this repo: against oracle/eigen_shim, see
// tests/test_dropin.py and INTEGRATION.md).
#pragma once
#include <memory>

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
