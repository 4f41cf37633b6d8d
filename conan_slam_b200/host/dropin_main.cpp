// dropin_main.cpp — drop-in demonstration against THE REFERENCE'S OWN HEADERS: the driver below
// talks only to `std::shared_ptr<Slam>` (slam/include/slam.h), uses the reference's own inline
// simulator helpers (computeSWA, vehicleModel, getObservations — inherited from `Slam`), and selects
// the implementation with one line:  new EKF(LM, WP)  (reference, CPU)  or  new EKFGpu(LM, WP).
// Same call sequence as test/main.cpp:132-200 with the two noise switches off (reproducible).
// Built by oracle/Makefile (target _ref) against oracle/eigen_shim; needs /root/reference at
// build time only.
//   dropin_main gpu|ref [max_steps] [trace_file]            EKF loop of test/main.cpp:132-200
//   dropin_main pfgpu|pfref [particles] [dump_file]         the PER-PARTICLE loops of test/main.cpp:279-327
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>

#include "EKF.h"
#include "PF.h"
#include "slam_gpu_eigen.hpp"

// The particle-filter loops of test/main.cpp:204-335, written against std::shared_ptr<Slam> with the
// reference's per-particle virtuals, run with `new PF(LM, WP)` (reference, CPU, FP32) or `new PFGpu(LM, WP)`.
// The driver's own table bookkeeping is bypassed (PF.cpp:204-213 never registers a landmark): the
// observations are handed over as "new" / "known" directly so that sampleProposal and featureUpdate DO run.
static int run_pf(bool use_gpu, int n, const char* dump, const Eigen::MatrixXf& LM, const Eigen::MatrixXf& WP) {
    std::shared_ptr<Slam> pf;
    if (use_gpu) pf.reset(new PFGpu(LM, WP));
    else pf.reset(new PF(LM, WP));
    pf->mNumParticles = n;
    auto particles = pf->initializeParticles(pf->mNumParticles);
    Eigen::MatrixXf Q = Eigen::MatrixXf::Zero(2, 2), R = Eigen::MatrixXf::Zero(2, 2);
    Q(0, 0) = pf->mSigmaV * pf->mSigmaV;
    Q(1, 1) = pf->mSigmaSWA * pf->mSigmaSWA;
    R(0, 0) = pf->mSigmaR * pf->mSigmaR;
    R(1, 1) = pf->mSigmaB * pf->mSigmaB;
    const Eigen::MatrixXf QE = 2 * Q, RE = 2 * R;  // test/main.cpp:244-245
    const float dt = pf->mDtControls;
    auto control = [&](int k0) {  // test/main.cpp:279-286, six control steps
        for (int k = k0; k < k0 + 6; k++) {
            const float swa = 0.02F * std::sin(0.3F * (float)k), phi = 0.0015F * (float)(k + 1);
            for (int pl = 0; pl < pf->mNumParticles; pl++) {
                pf->predict(particles[pl], pf->mVelocity, swa, QE, pf->mWheelBase, dt);
                pf->observeHeading(particles[pl], phi, pf->mSwitchHeadingKnown);
            }
        }
    };
    control(0);
    // first scan: three landmarks, all new -> the pose is sampled on the HOST by the driver (main.cpp:316-327)
    Eigen::MatrixXf ZN = Eigen::MatrixXf::Zero(2, 3);
    ZN(0, 0) = 410.0F; ZN(1, 0) = 0.31F;
    ZN(0, 1) = 930.0F; ZN(1, 1) = -0.52F;
    ZN(0, 2) = 1500.0F; ZN(1, 2) = 0.08F;
    for (int i = 0; i < pf->mNumParticles; i++) {
        particles[i].X = pf->multivariateNormalGaussianDistribution(particles[i].X, particles[i].P, 1);
        particles[i].P = Eigen::MatrixXf::Zero(3, 3);
        pf->addOneNewFeature(particles[i], ZN, RE);
    }
    const int nf1 = particles[0].XF.cols();  // main.cpp:299
    control(6);
    // second scan: the same three landmarks, now known (main.cpp:303-311 without the resampling)
    Eigen::MatrixXf ZF = Eigen::MatrixXf::Zero(2, 3);
    ZF(0, 0) = 401.0F; ZF(1, 0) = 0.305F;
    ZF(0, 1) = 921.5F; ZF(1, 1) = -0.533F;
    ZF(0, 2) = 1490.2F; ZF(1, 2) = 0.071F;
    Eigen::VectorXi IDF = Eigen::VectorXi::Zero(3);
    IDF(0) = 1; IDF(1) = 2; IDF(2) = 3;
    for (int i = 0; i < pf->mNumParticles; i++) {
        pf->sampleProposal(particles[i], ZF, IDF, RE);
        pf->featureUpdate(particles[i], ZF, IDF, RE);
    }
    Eigen::MatrixXf ZN2 = Eigen::MatrixXf::Zero(2, 1);
    ZN2(0, 0) = 700.0F; ZN2(1, 0) = 1.1F;
    for (int i = 0; i < pf->mNumParticles; i++) pf->addOneNewFeature(particles[i], ZN2, RE);
    const Eigen::VectorXf xs = pf->extractStatesFromParticles(particles);  // inline host helper, slam.h:493-511
    const int nf = particles[0].XF.cols();
    std::printf("%s: %d particles, features %d -> %d, extracted pose %.5f %.5f %.6f\n", use_gpu ? "pfgpu" : "pfref", n, nf1,
                nf, xs(0), xs(1), xs(2));
    if (dump) {
        FILE* f = fopen(dump, "wb");
        const float hdr[2] = {(float)n, (float)nf};
        fwrite(hdr, sizeof(float), 2, f);
        for (int i = 0; i < n; i++) {
            const Slam::Particle_t& p = particles[i];
            std::vector<float> rec;
            rec.push_back(p.w);
            for (int a = 0; a < 3; a++) rec.push_back(p.X(a));
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++) rec.push_back(p.P(a, b));
            for (int k = 0; k < nf; k++) { rec.push_back(p.XF(0, k)); rec.push_back(p.XF(1, k)); }
            for (int k = 0; k < nf; k++)
                for (int a = 0; a < 2; a++)
                    for (int b = 0; b < 2; b++) rec.push_back(p.PF[(size_t)k](a, b));
            fwrite(rec.data(), sizeof(float), rec.size(), f);
        }
        fclose(f);
    }
    if (use_gpu) {  // the reference's resampleParticles indexes particles[Keep - 1] with Keep = 0 (UB): ours only
        pf->resampleParticles(particles, pf->mNumParticles, true);
        float wsum = 0.0F;
        for (auto& p : particles) wsum += p.w;
        std::printf("pfgpu: resampled, sum of weights %.6f, features per particle %d\n", wsum, (int)particles[0].XF.cols());
    }
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && (!strcmp(argv[1], "pfgpu") || !strcmp(argv[1], "pfref"))) {
        extern const float kDropinLm[2][30];
        extern const float kDropinWp[2][5];
        Eigen::MatrixXf LM = Eigen::MatrixXf::Zero(2, 30), WP = Eigen::MatrixXf::Zero(2, 5);
        for (int i = 0; i < 30; i++) { LM(0, i) = kDropinLm[0][i]; LM(1, i) = kDropinLm[1][i]; }
        for (int i = 0; i < 5; i++) { WP(0, i) = kDropinWp[0][i]; WP(1, i) = kDropinWp[1][i]; }
        return run_pf(!strcmp(argv[1], "pfgpu"), argc > 2 ? atoi(argv[2]) : 64, argc > 3 ? argv[3] : nullptr, LM, WP);
    }
    const bool use_gpu = argc > 1 && !strcmp(argv[1], "gpu");
    const int max_steps = argc > 2 ? atoi(argv[2]) : (1 << 30);
    FILE* tf = argc > 3 ? fopen(argv[3], "wb") : nullptr;
    // the map and waypoints come from the tape generator's tables so this file holds no copy of them
    extern const float kDropinLm[2][30];
    extern const float kDropinWp[2][5];
    Eigen::MatrixXf LM = Eigen::MatrixXf::Zero(2, 30), WP = Eigen::MatrixXf::Zero(2, 5);
    for (int i = 0; i < 30; i++) { LM(0, i) = kDropinLm[0][i]; LM(1, i) = kDropinLm[1][i]; }
    for (int i = 0; i < 5; i++) { WP(0, i) = kDropinWp[0][i]; WP(1, i) = kDropinWp[1][i]; }

    std::shared_ptr<Slam> ekfSlam;
    if (use_gpu) ekfSlam.reset(new EKFGpu(LM, WP));
    else ekfSlam.reset(new EKF(LM, WP));
    ekfSlam->mSwitchControlNoise = false;
    ekfSlam->mSwitchSensorNoise = false;

    Eigen::MatrixXf Q = Eigen::MatrixXf::Zero(2, 2), R = Eigen::MatrixXf::Zero(2, 2);
    Q(0, 0) = ekfSlam->mSigmaV * ekfSlam->mSigmaV;
    Q(1, 1) = ekfSlam->mSigmaSWA * ekfSlam->mSigmaSWA;
    R(0, 0) = ekfSlam->mSigmaR * ekfSlam->mSigmaR;
    R(1, 1) = ekfSlam->mSigmaB * ekfSlam->mSigmaB;
    Eigen::VectorXf XTrue = Eigen::VectorXf::Zero(3), X = Eigen::VectorXf::Zero(3);
    Eigen::MatrixXf P = Eigen::MatrixXf::Zero(3, 3);
    const double dt = ekfSlam->mDtControls;
    double dtsum = 0.0;
    Eigen::VectorXi FeatureTag = Eigen::VectorXi::Zero(30);
    for (int i = 0; i < 30; i++) FeatureTag(i) = i + 1;
    int iwp = 1, steps = 0;
    float swa = 0.0F;
    Eigen::MatrixXf QE = 2 * Q, RE = 8 * R;
    while (iwp <= ekfSlam->getWayPoints().cols() && iwp > 0 && steps < max_steps) {
        steps++;
        ekfSlam->computeSWA(XTrue, ekfSlam->getWayPoints(), iwp, ekfSlam->mAtWaypoint, swa, ekfSlam->mRateSWA,
                            ekfSlam->mMaxSWA, dt);
        ekfSlam->vehicleModel(XTrue, ekfSlam->mVelocity, swa, ekfSlam->mWheelBase, dt);
        auto cn = ekfSlam->addControlNoise(ekfSlam->mVelocity, swa, Q, ekfSlam->mSwitchControlNoise);
        ekfSlam->predict(X, P, cn.v, cn.swa, QE, ekfSlam->mWheelBase, dt);
        ekfSlam->observeHeading(X, P, XTrue(2), ekfSlam->mSwitchHeadingKnown);
        dtsum = dtsum + dt;
        if (dtsum >= ekfSlam->mDtObserve) {
            dtsum = 0.0;
            auto obs = ekfSlam->getObservations(XTrue, ekfSlam->getLandMarks(), FeatureTag, ekfSlam->mMaxRange);
            ekfSlam->addObservationNoise(obs.Z, R, ekfSlam->mSwitchSensorNoise);
            if (obs.Z.size() > 0) {
                auto a = ekfSlam->dataAssociateTable(X, obs.Z, obs.idf, ekfSlam->mTABLE);
                ekfSlam->update(X, P, a.ZF, RE, a.idf, ekfSlam->mSwitchBatchUpdate);
                ekfSlam->augment(X, P, a.ZN, RE);
            }
        }
        if (tf) {
            const float rec[4] = {X(0), X(1), X(2), (float)X.rows()};
            fwrite(rec, sizeof(float), 4, tf);
        }
    }
    std::printf("%s: %d steps n=%d X0..2 = %.6f %.6f %.6f\n", use_gpu ? "gpu" : "ref", steps, (int)X.rows(), X(0), X(1),
                X(2));
    if (tf) {
        fwrite(X.data(), sizeof(float), (size_t)X.rows(), tf);
        fclose(tf);
    }
    return 0;
}
