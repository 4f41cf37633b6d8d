// oracle/oracle_bench.cpp — TEST/BENCH INFRASTRUCTURE (see slam_oracle.hpp header).
// CPU-baseline timing loops for bench.py's cpu_baseline / --impl reference legs: the
// reference's DENSE algorithm (slam.h:235-266 choleskyUpdate with a materialised H and an
// n x n temporary; EKF.cpp:131-144 computeAssociation with a dense H*P*H^T per pair; PF.cpp
// per-particle AoS steps), timed on a BOUNDED column slab of the full-size problem and scaled
// by n/c — the loops stream P column by column, so a slab of c columns costs exactly c/n of
// the full pass.  Threads split the slab's columns (the reference itself is single-threaded,
// CMakeLists.txt:73-77; threads > 1 is the generous baseline).
#include <chrono>
#include <random>
#include <thread>

#include "slam_oracle.hpp"

using namespace oracle;
using clk = std::chrono::steady_clock;

namespace {

template <class F>
void parallel_cols(int c, int threads, F f) {
    if (threads <= 1) {
        f(0, c, 0);
        return;
    }
    std::vector<std::thread> ts;
    for (int t = 0; t < threads; t++) {
        const int a = (int)((long long)c * t / threads), b = (int)((long long)c * (t + 1) / threads);
        ts.emplace_back([=] { f(a, b, t); });
    }
    for (auto& t : ts) t.join();
}

}  // namespace

extern "C" {

// One dense choleskyUpdate (rank r) restricted to a slab of c columns of an n x n covariance.
// Returns seconds for the slab (best of reps); full-size cost = seconds * n / c.
double orc_bench_dense_update_slab(int n, int c, int r, int threads, int reps) {
    if (c > n) c = n;
    std::vector<double> P((size_t)n * c), WWt((size_t)n * c);
    std::vector<double> Ht((size_t)c * r), W1((size_t)n * r);
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> ud(-1.0, 1.0);
    for (auto& v : P) v = ud(rng);
    for (auto& v : Ht) v = ud(rng);
    for (auto& v : W1) v = ud(rng) * 1e-3;
    double best = 1e300;
    for (int rep = 0; rep < reps; rep++) {
        const auto t0 = clk::now();
        // (1) PHT = P * H^T  (slam.h:243): j-k-i order, every column of the slab is streamed once
        std::vector<std::vector<double>> part(std::max(1, threads), std::vector<double>((size_t)n * r, 0.0));
        parallel_cols(c, threads, [&](int a, int b, int t) {
            double* pht = part[t].data();
            for (int j = 0; j < r; j++)
                for (int k = a; k < b; k++) {
                    const double h = Ht[(size_t)k * r + j];
                    const double* pk = &P[(size_t)k * n];
                    double* out = pht + (size_t)j * n;
                    for (int i = 0; i < n; i++) out[i] += pk[i] * h;
                }
        });
        // (4) the n x n temporary W1 * W1^T (slam.h:260), slab columns
        parallel_cols(c, threads, [&](int a, int b, int) {
            for (int j = a; j < b; j++) {
                double* col = &WWt[(size_t)j * n];
                for (int i = 0; i < n; i++) col[i] = 0.0;
                for (int k = 0; k < r; k++) {
                    const double w = W1[(size_t)k * n + (j % n)];
                    const double* wk = &W1[(size_t)k * n];
                    for (int i = 0; i < n; i++) col[i] += wk[i] * w;
                }
            }
        });
        // (5) P = P - temporary
        parallel_cols(c, threads, [&](int a, int b, int) {
            for (int j = a; j < b; j++) {
                double* pc = &P[(size_t)j * n];
                const double* wc = &WWt[(size_t)j * n];
                for (int i = 0; i < n; i++) pc[i] = pc[i] - wc[i];
            }
        });
        const double dt = std::chrono::duration<double>(clk::now() - t0).count();
        best = std::min(best, dt);
        volatile double sink = part[0][0] + P[0];
        (void)sink;
    }
    return best;
}

// One dense computeAssociation pair (EKF.cpp:140: S = H * P * H^T + R with a 2 x n H) on a slab
// of c columns of P.  Returns seconds for the slab; full-size per-pair cost = seconds * n / c.
double orc_bench_dense_gate_pair_slab(int n, int c, int threads, int reps) {
    if (c > n) c = n;
    std::vector<double> P((size_t)n * c), H((size_t)2 * n), HP((size_t)2 * c);
    std::mt19937_64 rng(9);
    std::uniform_real_distribution<double> ud(-1.0, 1.0);
    for (auto& v : P) v = ud(rng);
    for (int k = 0; k < 5; k++) {
        H[2 * (size_t)(k * (n / 5))] = ud(rng);
        H[2 * (size_t)(k * (n / 5)) + 1] = ud(rng);
    }
    double best = 1e300;
    for (int rep = 0; rep < reps; rep++) {
        const auto t0 = clk::now();
        parallel_cols(c, threads, [&](int a, int b, int) {
            for (int j = a; j < b; j++) {  // HP(:,j) = H * P(:,j), dense over all n rows
                const double* pj = &P[(size_t)j * n];
                double s0 = 0.0, s1 = 0.0;
                for (int k = 0; k < n; k++) {
                    s0 += H[2 * (size_t)k] * pj[k];
                    s1 += H[2 * (size_t)k + 1] * pj[k];
                }
                HP[2 * (size_t)j] = s0;
                HP[2 * (size_t)j + 1] = s1;
            }
        });
        const double dt = std::chrono::duration<double>(clk::now() - t0).count();
        best = std::min(best, dt);
        volatile double sink = HP[0];
        (void)sink;
    }
    return best;
}

// Reference-style PF step (AoS particles, per-call heap allocation as in PF.cpp) on a bounded
// number of particles: predict + heading + sampleProposal + featureUpdate for m_obs known
// associations, then one stratified resample with a deep copy of the survivors.
// Returns seconds per particle-step.
double orc_bench_pf_step(int num_particles, int num_features, int m_obs, int threads, unsigned flags) {
    typedef double T;
    std::vector<Particle<T>> ps = pf_initialize_particles<T>(num_particles);
    std::mt19937_64 rng(11);
    std::normal_distribution<double> nd(0.0, 1.0);
    Mat<T> Q(2, 2), R(2, 2);
    Q(0, 0) = 2 * 0.09; Q(1, 1) = 2 * 3.0461741978670857e-4;
    R(0, 0) = 2 * 0.01; R(1, 1) = 2 * 3.0461741978670857e-4;
    Mat<T> Z0(2, num_features);
    for (int f = 0; f < num_features; f++) {
        Z0(0, f) = 200.0 + 1500.0 * (double)f / num_features;
        Z0(1, f) = -1.2 + 2.4 * (double)f / num_features;
    }
    for (auto& p : ps) {
        for (int k = 0; k < 6; k++) {
            pf_predict<T>(p, 83.33, 0.02, Q, 73.0, 0.01);
            pf_observe_heading<T>(p, 0.001, true);
        }
        pf_add_new_features<T>(p, Z0, R);
    }
    Mat<T> Z(2, m_obs);
    std::vector<int> idf(m_obs);
    for (int k = 0; k < m_obs; k++) {
        idf[k] = 1 + k * (num_features / m_obs);
        Z(0, k) = Z0(0, idf[k] - 1) + 0.002;
        Z(1, k) = Z0(1, idf[k] - 1) + 1e-6;
    }
    std::vector<double> xi(3 * (size_t)num_particles), u(num_particles);
    for (auto& v : xi) v = nd(rng);
    for (auto& v : u) v = 0.3 * nd(rng);
    const auto t0 = clk::now();
    parallel_cols(num_particles, threads, [&](int a, int b, int) {
        for (int i = a; i < b; i++) {
            pf_predict<T>(ps[i], 83.33, 0.02, Q, 73.0, 0.01);
            pf_observe_heading<T>(ps[i], 0.0012, true);
            pf_sample_proposal<T>(ps[i], Z, idf, R, &xi[3 * (size_t)i], flags);
            pf_feature_update<T>(ps[i], Z, idf, R, flags);
        }
    });
    pf_resample_particles<T>(ps, u, (T)(num_particles + 1), true, flags | FLAG_Q10_SEARCH);
    const double dt = std::chrono::duration<double>(clk::now() - t0).count();
    return dt / num_particles;
}

int orc_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
