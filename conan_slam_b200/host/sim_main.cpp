// sim_main.cpp — the reference's simulation driver (test/main.cpp:15-201, EKF branch) replayed
// line for line through the C++ adaptor (slam_gpu.hpp) and the C ABI: same map, waypoints,
// control loop, known-association table logic, update / augment calls.  Config 1 of
// BASELINE.json ("bit-match run": noise switches off, FP64).
//
//   sim_main [--flags N] [--max-steps N] [--trace FILE] [--gated] [--print-every K] [--fused] [--no-cov-writeback]
// --fused: the control steps between two observations in ONE launch (cslam_ekf_control_steps) and the observation
// step (joint update + augmentation) in ONE launch (cslam_ekf_observe_step); --no-cov-writeback: P stays on the
// device (the driver only prints X, test/main.cpp:134-137), the host copy of P is not refreshed after each call.
//
// --trace writes, per control step, (x, y, phi, n) of the ESTIMATE as 4 doubles, then the final
// full state X; tests/test_host_cpp.py compares it with the CPU oracle replaying the same tape.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "slam_gpu.hpp"

using namespace cslam_host;

namespace {
const double kPi = 3.14159265358979323846;

// test/main.cpp:24-54 and :67-76 — the reference stores the map as FP32 literals
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

double pi2Pi(double a) {  // slam.h:816-829
    a = std::fmod(a, 2 * kPi);
    if (a > kPi) a = a - 2.0 * kPi;
    if (a < -kPi) a = a + 2 * kPi;
    return a;
}
int signumInt(int x) { return (0 < x) - (x < 0); }  // slam.h:924-928 instantiated as signum<int>

// slam.h:279-332 (signum<int>(float) truncates its argument first)
void computeSWA(const DVec& X, const DMat& WP, int& iwp, double minD, double& swa, double rateSWA, double maxSWA,
                double dt) {
    if (WP.cols() <= 0) return;
    double cx = WP(0, iwp - 1), cy = WP(1, iwp - 1);
    const double d2 = (cx - X(0)) * (cx - X(0)) + (cy - X(1)) * (cy - X(1));
    if (d2 < minD * minD) {
        iwp = iwp + 1;
        if (iwp > WP.cols()) { iwp = 0; return; }
        cx = WP(0, iwp - 1);
        cy = WP(1, iwp - 1);
    }
    double deltaG = pi2Pi(std::atan2(cy - X(1), cx - X(0)) - X(2) - swa);
    const double maxDelta = rateSWA * dt;
    if (std::fabs(deltaG) > maxDelta) deltaG = maxDelta * signumInt(static_cast<int>(deltaG));
    swa = swa + deltaG;
    if (std::fabs(swa) > maxSWA) swa = signumInt(static_cast<int>(swa)) * maxSWA;
}
// slam.h:952-966
void vehicleModel(DVec& X, double v, double swa, double wb, double dt) {
    const double x0 = X(0), x1 = X(1), x2 = X(2);
    X(0) = x0 + v * dt * std::cos(swa + x2);
    X(1) = x1 + v * dt * std::sin(swa + x2);
    X(2) = pi2Pi(x2 + v * dt * std::sin(swa) / wb);
}
// slam.h:575-582 -> :608-683 -> :339-368
void getObservations(const DVec& X, const DMat& LM, double rmax, DMat& Z, std::vector<int>& tags) {
    std::vector<int> vis;
    const double phi = X(2);
    for (int i = 0; i < LM.cols(); i++) {
        const double dx = LM(0, i) - X(0), dy = LM(1, i) - X(1);
        if ((std::fabs(dx) < rmax && std::fabs(dy) < rmax) && ((dx * std::cos(phi) + dy * std::sin(phi)) > 0.0) &&
            ((dx * dx + dy * dy) < rmax * rmax))
            vis.push_back(i);
    }
    Z.resize(vis.empty() ? 0 : 2, (int)vis.size());
    tags.clear();
    for (size_t k = 0; k < vis.size(); k++) {
        const double dx = LM(0, vis[k]) - X(0), dy = LM(1, vis[k]) - X(1);
        Z(0, (int)k) = std::sqrt(dx * dx + dy * dy);
        Z(1, (int)k) = std::atan2(dy, dx) - X(2);
        tags.push_back(vis[k] + 1);
    }
}
}  // namespace

int main(int argc, char** argv) {
    unsigned flags = CSLAM_FLAG_REF_LITERAL;
    int max_steps = 1 << 30, print_every = 2000;
    bool gated = false, fused = false, no_writeback = false;
    std::string trace;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--flags") && i + 1 < argc) flags = (unsigned)atoi(argv[++i]);
        else if (!strcmp(argv[i], "--max-steps") && i + 1 < argc) max_steps = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--trace") && i + 1 < argc) trace = argv[++i];
        else if (!strcmp(argv[i], "--print-every") && i + 1 < argc) print_every = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--gated")) gated = true;
        else if (!strcmp(argv[i], "--fused")) fused = true;  // control steps between observations in one call
        else if (!strcmp(argv[i], "--no-cov-writeback")) no_writeback = true;
    }
    DMat LM(2, 30), WP(2, 5);
    for (int i = 0; i < 30; i++) { LM(0, i) = kLm1[i]; LM(1, i) = kLm2[i]; }
    for (int i = 0; i < 5; i++) { WP(0, i) = kWp1[i]; WP(1, i) = kWp2[i]; }

    // std::shared_ptr<Slam> ekfSlam(new EKF(LM, WP));            test/main.cpp:89
    EkfGpu ekfSlam(LM.cols(), LM.cols(), 0, flags);
    ekfSlam.mSwitchAssociationKnown = !gated;
    if (no_writeback) ekfSlam.p_writeback_max = 0;
    const double sigmaV = 0.3F, sigmaSWA = (float)(1.0F * kPi / 180.0F);
    const double sigmaR = 0.1F, sigmaB = (float)(1.0F * kPi / 180.0F);
    const double velocity = 83.33F, wheelBase = 73.0F, maxRange = 2000.0F, atWaypoint = 1.0F;
    const double maxSWA = (float)(kPi / 4.0F), rateSWA = (float)(70.0F * kPi / 180.0F);
    DMat Q(2, 2), R(2, 2);
    Q(0, 0) = sigmaV * sigmaV; Q(1, 1) = sigmaSWA * sigmaSWA;   // test/main.cpp:93-97
    R(0, 0) = sigmaR * sigmaR; R(1, 1) = sigmaB * sigmaB;       // test/main.cpp:99-103
    DVec XTrue(3), X(3);
    DMat P(3, 3);
    const double dt = 0.01, dtObserve = 5.058F * 0.01;
    double dtsum = 0.0;
    std::vector<int> FeatureTag(30);
    for (int i = 0; i < 30; i++) FeatureTag[i] = i + 1;
    int iwp = 1;
    double swa = 0.0;
    DMat QE = Q, RE = R;  // mSwitchInflateNoise (test/main.cpp:125-129)
    for (auto& v : QE.a) v *= 2;
    for (auto& v : RE.a) v *= 8;

    FILE* tf = trace.empty() ? nullptr : fopen(trace.c_str(), "wb");
    const auto t_start = std::chrono::steady_clock::now();
    std::vector<double> pend_v, pend_s, pend_p;
    int indexlooper = 0;
    while (iwp <= WP.cols() && iwp > 0 && indexlooper < max_steps) {
        ++indexlooper;
        if (print_every > 0 && indexlooper % print_every == 0)
            std::printf("indexlooper\t%d\tX\t%.9g %.9g %.9g\tn=%d\n", indexlooper, X(0), X(1), X(2), X.rows());
        computeSWA(XTrue, WP, iwp, atWaypoint, swa, rateSWA, maxSWA, dt);          // test/main.cpp:140
        vehicleModel(XTrue, velocity, swa, wheelBase, dt);                         // :156
        const double vn = velocity, swan = swa;                                    // :162 (control noise off)
        dtsum = dtsum + dt;
        const bool observe = dtsum >= dtObserve;                                   // :172
        if (fused) {
            // the controls do not depend on the filter: buffer them and hand all steps up to the next
            // observation to ONE call (cslam_ekf_control_steps; a single launch on this 30-landmark map)
            pend_v.push_back(vn); pend_s.push_back(swan); pend_p.push_back(XTrue(2));
            const bool last = !(iwp <= WP.cols() && iwp > 0 && indexlooper < max_steps);
            if (observe || last || pend_v.size() >= 64) {
                std::vector<double> tr;
                ekfSlam.controlSteps(X, P, pend_v, pend_s, pend_p, ekfSlam.mSwitchHeadingKnown, QE, wheelBase, dt, &tr);
                if (tf) {
                    for (size_t k = 0; k + 1 < pend_v.size(); k++) {  // the last step is recorded below
                        const double rec[4] = {tr[3 * k], tr[3 * k + 1], tr[3 * k + 2], (double)X.rows()};
                        fwrite(rec, sizeof(double), 4, tf);
                    }
                }
                pend_v.clear(); pend_s.clear(); pend_p.clear();
            } else {
                continue;  // nothing to record yet: this step's pose comes back with the batch
            }
        } else {
            ekfSlam.predict(X, P, vn, swan, QE, wheelBase, dt);                    // :165
            ekfSlam.observeHeading(X, P, XTrue(2), ekfSlam.mSwitchHeadingKnown);   // :168
        }
        if (observe) {
            dtsum = 0.0;
            DMat Z;
            std::vector<int> visible;
            getObservations(XTrue, LM, maxRange, Z, visible);                      // :177 (sensor noise off)
            if (Z.size() > 0) {
                if (ekfSlam.mSwitchAssociationKnown) {
                    auto a = ekfSlam.dataAssociateTable(X, Z, visible, ekfSlam.mTABLE);   // :186
                    if (fused) {  // :188-189 as one single-CTA launch (cslam_ekf_observe_step)
                        ekfSlam.observeStep(X, P, a.ZF, RE, a.idf, a.ZN, ekfSlam.mSwitchBatchUpdate);
                    } else {
                        ekfSlam.update(X, P, a.ZF, RE, a.idf, ekfSlam.mSwitchBatchUpdate);    // :188
                        ekfSlam.augment(X, P, a.ZN, RE);                                      // :189
                    }
                } else {
                    auto a = ekfSlam.dataAssociate(X, P, Z, RE, ekfSlam.mGateReject, ekfSlam.mGateAugment);  // :194
                    ekfSlam.update(X, P, a.ZF, RE, a.idf, ekfSlam.mSwitchBatchUpdate);    // :195
                    ekfSlam.augment(X, P, a.ZN, RE);                                      // :196
                }
            }
        }
        if (tf) {
            const double rec[4] = {X(0), X(1), X(2), (double)X.rows()};
            fwrite(rec, sizeof(double), 4, tf);
        }
    }
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    std::printf("done: %d control steps, n=%d, skipped updates=%d, loop wall time %.3f s\nX", indexlooper, X.rows(),
                ekfSlam.skippedUpdates(), wall);
    for (int i = 0; i < X.rows(); i++) std::printf(" %.12g", X(i));
    std::printf("\n");
    if (tf) {
        fwrite(X.a.data(), sizeof(double), X.a.size(), tf);
        fclose(tf);
    }
    return EXIT_SUCCESS;
}
