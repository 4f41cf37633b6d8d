// dropin_main.cpp — drop-in demonstration against THE REFERENCE'S OWN HEADERS: the driver below
// talks only to `std::shared_ptr<Slam>` (slam/include/slam.h), uses the reference's own inline
// simulator helpers (computeSWA, vehicleModel, getObservations — inherited from `Slam`), and selects
// the implementation with one line:  new EKF(LM, WP)  (reference, CPU)  or  new EKFGpu(LM, WP).
// Same call sequence as test/main.cpp:132-200 with the two noise switches off (reproducible).
// Built by oracle/Makefile (target _ref) against oracle/eigen_shim; needs /root/reference at
// build time only.
//   dropin_main gpu|ref [max_steps] [trace_file]
#include <cstdio>
#include <cstring>
#include <memory>

#include "EKF.h"
#include "slam_gpu_eigen.hpp"

int main(int argc, char** argv) {
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
