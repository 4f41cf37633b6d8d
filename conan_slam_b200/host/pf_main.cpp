// pf_main.cpp — the particle-filter half of test/main.cpp (:204-335) driven through the C++
// population-level adaptor (slam_gpu.hpp PfGpuT -> C ABI -> CUDA): a fixed synthetic sequence
// (6 control steps, pose sampling + map initialisation, then observation cycles with resampling)
// whose final weights / poses are dumped for tests/test_host_cpp.py, which replays the identical
// sequence through the Python mirror and the CPU oracle.
//   usage: pf_main --particles P --cycles C --flags F --out file.bin
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "slam_gpu.hpp"

using namespace cslam_host;

static double lcg_normalish(unsigned long long& s) {  // deterministic stand-in draw tape (inputs)
    double acc = 0.0;
    for (int k = 0; k < 4; k++) {
        s = s * 6364136223846793005ULL + 1442695040888963407ULL;
        acc += (double)((s >> 11) & 0xFFFFFFFFFFFFFULL) / 4503599627370496.0;
    }
    return (acc - 2.0) * 1.7320508075688772;
}

int main(int argc, char** argv) {
    int P = 512, cycles = 3;
    unsigned flags = CSLAM_FLAG_INTENDED;
    const char* out = nullptr;
    for (int i = 1; i + 1 < argc; i += 2) {
        if (!strcmp(argv[i], "--particles")) P = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--cycles")) cycles = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--flags")) flags = (unsigned)atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--out")) out = argv[i + 1];
    }
    const int NF = 3;
    PfGpu pf(NF, P, 8, 0, flags);
    DMat Q(2, 2), R(2, 2);
    Q(0, 0) = 2 * 0.3 * 0.3; Q(1, 1) = 2 * std::pow(M_PI / 180.0, 2);   // QE = 2Q  main.cpp:244
    R(0, 0) = 2 * 0.1 * 0.1; R(1, 1) = 2 * std::pow(M_PI / 180.0, 2);   // RE = 2R  main.cpp:245
    unsigned long long seed = 42;
    auto draws = [&](int count) { std::vector<double> v((size_t)count); for (auto& x : v) x = lcg_normalish(seed); return v; };
    auto controls = [&](int c0) {
        for (int c = 0; c < 6; c++) {
            pf.predict(83.33, 0.02 * std::sin(0.3 * (c0 + c)), Q, 73.0, 0.01);
            pf.observeHeading(0.001 * (c0 + c), true);
        }
    };
    controls(0);
    pf.samplePose(draws(3 * P));
    DMat ZN(2, NF);
    const double zr[NF] = {400.0, 900.0, 650.0}, zb[NF] = {0.3, -0.7, 0.05};
    for (int k = 0; k < NF; k++) { ZN(0, k) = zr[k]; ZN(1, k) = zb[k]; }
    pf.addOneNewFeature(ZN, R);
    std::vector<int> ids = {1, 3};
    double neff = 0.0;
    std::vector<int32_t> keep;
    for (int c = 0; c < cycles; c++) {
        controls(6 * (c + 1));
        DMat ZF(2, 2);
        ZF(0, 0) = zr[0] - 5.0 * (c + 1); ZF(1, 0) = zb[0] + 0.002 * (c + 1);
        ZF(0, 1) = zr[2] - 5.0 * (c + 1); ZF(1, 1) = zb[2] - 0.001 * (c + 1);
        pf.sampleProposal(ZF, ids, R, draws(3 * P));
        pf.featureUpdate(ZF, ids, R);
        std::vector<double> u = draws(P);
        for (auto& x : u) x *= 0.3;
        keep = pf.resampleParticles(1e300, u, true, &neff);
    }
    int idx = 0;
    DVec best = pf.extractStatesFromParticles(&idx);
    std::vector<double> w = pf.weights(), x = pf.poses();
    printf("particles=%d features=%d neff=%.17g best=%d skipped=%d\n", P, pf.numFeatures(), neff, idx, pf.skippedUpdates());
    if (out) {
        FILE* f = fopen(out, "wb");
        if (!f) return 2;
        fwrite(w.data(), sizeof(double), w.size(), f);
        fwrite(x.data(), sizeof(double), x.size(), f);
        std::vector<double> k(keep.begin(), keep.end());
        fwrite(k.data(), sizeof(double), k.size(), f);
        fwrite(&neff, sizeof(double), 1, f);
        fclose(f);
    }
    return 0;
}
