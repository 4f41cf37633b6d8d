// oracle/slam_oracle.hpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement (dependency-free C++17, templated on the scalar) of the filter
// arithmetic of mfkiwl/conan-slam: slam/src/EKF.cpp, slam/src/PF.cpp and the shared
// algebra / simulator helpers in slam/include/slam.h.  Every function cites the
// reference file:line it follows (paths relative to /root/reference).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build, link, import or execute this file.  The product (conan_slam_b200/)
// never does.
//
// PARITY PINNING: the reference ships no tests, golden vectors or known answers
//  and depends on Eigen 3.4.0 / Boost 1.84.0 (conanfile.py:53-55),
// neither of which exists in this image.  The oracle is therefore pinned two ways:
//   (1) oracle/_ref: the reference's OWN sources (EKF.cpp, PF.cpp, slam.h, untouched,
//       compiled where they lie) built against the minimal Eigen-API shim in
//       oracle/eigen_shim/ — see oracle/Makefile and tests/test_oracle_vs_ref.py;
//   (2) an independent numpy restatement (tests/np_ref.py).
// What stays unpinned: Eigen's own operation order inside GEMM/LLT/PartialPivLU and
// Boost's normal variates (third-party, not vendored).  Those differ from any
// restatement only at rounding level (FP64: ~1e-15 relative, far inside the 1e-9
// budget); random draws are INPUTS to both oracle and GPU path.
//
// Quirk flags.  0 == REF_LITERAL reproduces the reference
// bit-for-bit in structure; each bit switches ONE quirk to the intended behaviour.
// The same bit values are used by include/cslam.h.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <vector>

namespace oracle {

enum : unsigned {
    FLAG_REF_LITERAL   = 0u,
    FLAG_Q1_METRIC_S   = 1u << 0,  // gain metric M = S = L L^T instead of literal L^T L (slam.h:250-260)
    FLAG_Q2_FULL_WIDTH = 1u << 1,  // predict cross-covariance over all columns (EKF.cpp:442-443 uses cols()-4)
    FLAG_Q5_RETURN_ZN  = 1u << 2,  // dataAssociate returns the new-feature set (EKF.cpp:308-315 shadows it)
    FLAG_Q9_METRIC_S   = 1u << 3,  // gaussEvaluate exponent uses S^-1 (PF.cpp:287-300 uses (L^T L)^-1)
    FLAG_Q10_SEARCH    = 1u << 4,  // resampling: Keep[c] = min{i: select[c] < cumW[i]} (PF.cpp:566-574)
    FLAG_INTENDED      = 0x1Fu
};

static constexpr double kPi = 3.14159265358979323846;  // std::_Pi_val is a double-valued pi (slam.h:66)

// ---------------------------------------------------------------------------------
// Dense column-major matrix (mirrors Eigen::Matrix<T,Dynamic,Dynamic> storage).
// ---------------------------------------------------------------------------------
template <class T>
struct Mat {
    int r = 0, c = 0;
    std::vector<T> a;
    Mat() = default;
    Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, T(0)) {}
    T& operator()(int i, int j) { return a[(size_t)j * r + i]; }
    const T& operator()(int i, int j) const { return a[(size_t)j * r + i]; }
    bool all_finite() const {
        for (const T& v : a)
            if (!std::isfinite(v)) return false;
        return true;
    }
    static Mat identity(int n) {
        Mat m(n, n);
        for (int i = 0; i < n; i++) m(i, i) = T(1);
        return m;
    }
};
template <class T>
using Vec = std::vector<T>;

template <class T>
Mat<T> matmul(const Mat<T>& A, const Mat<T>& B) {  // C = A*B, j-k-i order (column-major friendly)
    Mat<T> C(A.r, B.c);
    for (int j = 0; j < B.c; j++)
        for (int k = 0; k < A.c; k++) {
            const T b = B(k, j);
            const T* ak = &A.a[(size_t)k * A.r];
            T* cj = &C.a[(size_t)j * C.r];
            for (int i = 0; i < A.r; i++) cj[i] += ak[i] * b;
        }
    return C;
}
template <class T>
Mat<T> transpose(const Mat<T>& A) {
    Mat<T> B(A.c, A.r);
    for (int j = 0; j < A.c; j++)
        for (int i = 0; i < A.r; i++) B(j, i) = A(i, j);
    return B;
}
template <class T>
Mat<T> add(const Mat<T>& A, const Mat<T>& B) {
    Mat<T> C(A.r, A.c);
    for (size_t i = 0; i < A.a.size(); i++) C.a[i] = A.a[i] + B.a[i];
    return C;
}
template <class T>
Mat<T> sub(const Mat<T>& A, const Mat<T>& B) {
    Mat<T> C(A.r, A.c);
    for (size_t i = 0; i < A.a.size(); i++) C.a[i] = A.a[i] - B.a[i];
    return C;
}

// Eigen's dynamic-size inverse()/determinant() go through PartialPivLU (third-party,
// Eigen 3.4.0 src/LU/PartialPivLU.h).  Restated as textbook LU with row pivoting.
template <class T>
struct PartialPivLU {
    Mat<T> lu;
    std::vector<int> perm;
    int sign = 1;
    explicit PartialPivLU(const Mat<T>& M) : lu(M), perm(M.r) {
        const int n = M.r;
        for (int i = 0; i < n; i++) perm[i] = i;
        for (int k = 0; k < n; k++) {
            int piv = k;
            T best = std::abs(lu(k, k));
            for (int i = k + 1; i < n; i++)
                if (std::abs(lu(i, k)) > best) { best = std::abs(lu(i, k)); piv = i; }
            if (piv != k) {
                for (int j = 0; j < n; j++) std::swap(lu(k, j), lu(piv, j));
                std::swap(perm[k], perm[piv]);
                sign = -sign;
            }
            for (int i = k + 1; i < n; i++) {
                lu(i, k) = lu(i, k) / lu(k, k);
                const T f = lu(i, k);
                for (int j = k + 1; j < n; j++) lu(i, j) -= f * lu(k, j);
            }
        }
    }
    T determinant() const {
        T d = T(sign);
        for (int i = 0; i < lu.r; i++) d *= lu(i, i);
        return d;
    }
    Mat<T> inverse() const {
        const int n = lu.r;
        Mat<T> inv(n, n);
        for (int col = 0; col < n; col++) {
            std::vector<T> y(n);
            for (int i = 0; i < n; i++) {  // L y = P e_col
                T s = (perm[i] == col) ? T(1) : T(0);
                for (int k = 0; k < i; k++) s -= lu(i, k) * y[k];
                y[i] = s;
            }
            for (int i = n - 1; i >= 0; i--) {  // U x = y
                T s = y[i];
                for (int k = i + 1; k < n; k++) s -= lu(i, k) * inv(k, col);
                inv(i, col) = s / lu(i, i);
            }
        }
        return inv;
    }
};
template <class T>
Mat<T> inverse(const Mat<T>& M) {
    if (M.r == 0) return M;
    return PartialPivLU<T>(M).inverse();
}
template <class T>
T determinant(const Mat<T>& M) {
    if (M.r == 0) return T(1);
    return PartialPivLU<T>(M).determinant();
}

// slam.h:776-779
template <class T>
Mat<T> make_symmetric(const Mat<T>& P) {
    Mat<T> S(P.r, P.c);
    for (int j = 0; j < P.c; j++)
        for (int i = 0; i < P.r; i++) S(i, j) = (P(i, j) + P(j, i)) * T(0.5);
    return S;
}

// Cyclic Jacobi eigen-decomposition of a small symmetric matrix (stands in for
// Eigen::SelfAdjointEigenSolver in the fallback of slam.h:425-429; the factor it
// yields is not unique, so this branch is "structure only" — see header).
template <class T>
void jacobi_eigen(Mat<T> A, Mat<T>& V, Vec<T>& w) {
    const int n = A.r;
    V = Mat<T>::identity(n);
    for (int sweep = 0; sweep < 64; sweep++) {
        T off = 0;
        for (int p = 0; p < n; p++)
            for (int q = p + 1; q < n; q++) off += A(p, q) * A(p, q);
        if (!(off > T(0))) break;
        for (int p = 0; p < n; p++)
            for (int q = p + 1; q < n; q++) {
                if (A(p, q) == T(0)) continue;
                const T theta = (A(q, q) - A(p, p)) / (T(2) * A(p, q));
                const T t = (theta >= 0 ? T(1) : T(-1)) / (std::abs(theta) + std::sqrt(theta * theta + T(1)));
                const T cs = T(1) / std::sqrt(t * t + T(1)), sn = t * cs;
                for (int k = 0; k < n; k++) {
                    const T akp = A(k, p), akq = A(k, q);
                    A(k, p) = cs * akp - sn * akq;
                    A(k, q) = sn * akp + cs * akq;
                }
                for (int k = 0; k < n; k++) {
                    const T apk = A(p, k), aqk = A(q, k);
                    A(p, k) = cs * apk - sn * aqk;
                    A(q, k) = sn * apk + cs * aqk;
                }
                for (int k = 0; k < n; k++) {
                    const T vkp = V(k, p), vkq = V(k, q);
                    V(k, p) = cs * vkp - sn * vkq;
                    V(k, q) = sn * vkp + cs * vkq;
                }
            }
    }
    w.resize(n);
    for (int i = 0; i < n; i++) w[i] = A(i, i);
}

// slam.h:413-436 choleskyDecomposition.  *used_fallback reports the eigen-solver branch.
template <class T>
Mat<T> cholesky_decomposition(const Mat<T>& M, bool* used_fallback = nullptr) {
    const int n = M.r;
    Mat<T> L(n, n);
    bool ok = true;  // Eigen::LLT::info(): NumericalIssue when a pivot is <= 0 (reads lower triangle)
    for (int j = 0; j < n && ok; j++) {
        T d = M(j, j);
        for (int k = 0; k < j; k++) d -= L(j, k) * L(j, k);
        if (!(d > T(0))) { ok = false; break; }
        const T ljj = std::sqrt(d);
        L(j, j) = ljj;
        for (int i = j + 1; i < n; i++) {
            T s = M(i, j);
            for (int k = 0; k < j; k++) s -= L(i, k) * L(j, k);
            L(i, j) = s / ljj;
        }
    }
    if (used_fallback) *used_fallback = !ok;
    if (!ok) {  // slam.h:425-429
        Mat<T> V;
        Vec<T> w;
        jacobi_eigen(M, V, w);
        for (int j = 0; j < n; j++) {
            const T s = std::sqrt(w[j]);
            for (int i = 0; i < n; i++) L(i, j) = V(i, j) * s;
        }
    }
    if (!L.all_finite()) L = Mat<T>(n, n);  // slam.h:431-434
    return L;
}

// slam.h:816-829 pi2Pi.  The reference mixes float fmod with double-valued pi in the
// comparisons and corrections; T=double collapses to plain double arithmetic.
template <class T>
T pi2pi(T angle) {
    angle = std::fmod(angle, static_cast<T>(2 * kPi));
    if (static_cast<double>(angle) > kPi) angle = static_cast<T>(static_cast<double>(angle) - 2.0 * kPi);
    if (static_cast<double>(angle) < -kPi) angle = static_cast<T>(static_cast<double>(angle) + 2 * kPi);
    return angle;
}

// ---------------------------------------------------------------------------------
// Shared Kalman algebra
// ---------------------------------------------------------------------------------

// slam.h:235-266 choleskyUpdate — dense, literal (P*H^T over all n columns, n x n
// temporary for W1*W1^T).  Returns false when the update was numerically skipped
// (SCHOLINV not finite -> zero gain, slam.h:252-255).
template <class T>
bool cholesky_update(Vec<T>& X, Mat<T>& P, const Vec<T>& V, const Mat<T>& R, const Mat<T>& H, unsigned flags) {
    const int n = P.r, r = H.r;
    Mat<T> Ht = transpose(H);
    Mat<T> PHT = matmul(P, Ht);
    Mat<T> S = add(matmul(H, PHT), R);
    S = make_symmetric(S);
    Mat<T> SCHOL = cholesky_decomposition(S);
    Mat<T> SCHOLINV = inverse(SCHOL);
    bool applied = true;
    if (!SCHOLINV.all_finite()) {
        SCHOLINV = Mat<T>(r, r);
        applied = false;
    }
    // Q1: literal uses W1 = PHT*L^-1 (metric L^T L); intended uses W1 = PHT*L^-T (metric S).
    Mat<T> G = (flags & FLAG_Q1_METRIC_S) ? transpose(SCHOLINV) : SCHOLINV;
    Mat<T> W1 = matmul(PHT, G);
    Mat<T> W = matmul(W1, transpose(G));
    for (int i = 0; i < n; i++) {
        T s = 0;
        for (int k = 0; k < r; k++) s += W(i, k) * V[k];
        X[i] = X[i] + s;
    }
    Mat<T> WWt = matmul(W1, transpose(W1));  // the n x n temporary of slam.h:260
    for (size_t i = 0; i < P.a.size(); i++) P.a[i] = P.a[i] - WWt.a[i];
    return applied;
}

// slam.h:700-725 josephUpdate — dense, literal (two n^3 products).
template <class T>
void joseph_update_dense(Vec<T>& X, Mat<T>& P, const Vec<T>& V, const Mat<T>& R, const Mat<T>& H) {
    const int n = P.r;
    Mat<T> PHT = matmul(P, transpose(H));
    Mat<T> S = add(matmul(H, PHT), R);
    Mat<T> SI = make_symmetric(inverse(S));
    Mat<T> W = matmul(PHT, SI);
    for (int i = 0; i < n; i++) {
        T s = 0;
        for (int k = 0; k < W.c; k++) s += W(i, k) * V[k];
        X[i] = X[i] + s;
    }
    Mat<T> C = sub(Mat<T>::identity(n), matmul(W, H));
    Mat<T> Pn = add(matmul(matmul(C, P), transpose(C)), matmul(matmul(W, R), transpose(W)));
    const T tiny = static_cast<T>(std::numeric_limits<float>::min());  // slam.h:719 (Q3)
    for (int i = 0; i < n; i++) Pn(i, i) = Pn(i, i) + tiny;
    P = Pn;
}

// Same arithmetic as joseph_update_dense for the 1 x n selector H = e_2^T used by
// observeHeading (EKF.cpp:338-346), expanded term by term so it costs O(n^2):
//   (C P C^T)_ij = P_ij - W_i P_2j - (P_i2 - W_i P_22) W_j ;  + W_i R W_j ; + tiny on the diagonal.
// tests/test_oracle.py checks it against the dense form.
template <class T>
void joseph_update_heading(Vec<T>& X, Mat<T>& P, T v, T R) {
    const int n = P.r;
    const T S = P(2, 2) + R;
    const T SI = T(1) / S;
    Vec<T> W(n), row2(n), col2(n);
    for (int i = 0; i < n; i++) {
        W[i] = P(i, 2) * SI;
        row2[i] = P(2, i);
        col2[i] = P(i, 2);
    }
    for (int i = 0; i < n; i++) X[i] = X[i] + W[i] * v;
    const T p22 = P(2, 2);
    for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) {
            const T cp_ij = P(i, j) - W[i] * row2[j];
            const T cp_i2 = col2[i] - W[i] * p22;
            P(i, j) = (cp_ij - cp_i2 * W[j]) + (W[i] * R) * W[j];
        }
    const T tiny = static_cast<T>(std::numeric_limits<float>::min());
    for (int i = 0; i < n; i++) P(i, i) = P(i, i) + tiny;
}

// ---------------------------------------------------------------------------------
// EKF-SLAM (slam/src/EKF.cpp)
// ---------------------------------------------------------------------------------
template <class T>
struct ObserveModel {
    T z[2] = {0, 0};
    Mat<T> H;
};

// EKF.cpp:354-404 observeModel (dense 2 x n H; bearing NOT wrapped here).
template <class T>
ObserveModel<T> ekf_observe_model(const Vec<T>& X, int idf) {
    const int n = (int)X.size();
    const int fpos = 3 + idf * 2 - 1;
    ObserveModel<T> out;
    out.H = Mat<T>(2, n);
    if (n > 3) {
        const T dx = X[fpos - 1] - X[0];
        const T dy = X[fpos] - X[1];
        const T d2 = dx * dx + dy * dy;
        const T d = std::sqrt(d2);
        const T xd = dx / d, yd = dy / d, xd2 = dx / d2, yd2 = dy / d2;
        out.z[0] = d;
        out.z[1] = std::atan2(dy, dx) - X[2];
        out.H(0, 0) = -xd;  out.H(0, 1) = -yd;  out.H(0, 2) = T(0);
        out.H(1, 0) = yd2;  out.H(1, 1) = -xd2; out.H(1, 2) = T(-1);
        out.H(0, fpos - 1) = xd;   out.H(0, fpos) = yd;
        out.H(1, fpos - 1) = -yd2; out.H(1, fpos) = xd2;
    }
    return out;
}

// EKF.cpp:406-455 predict.  Covariance uses the OLD heading; state advanced last.
template <class T>
void ekf_predict(Vec<T>& X, Mat<T>& P, T v, T swa, const Mat<T>& Q, T wb, T dt, unsigned flags) {
    const T phi = X[2];
    Mat<T> Gv(3, 3), Gu(3, 2);
    Gv(0, 0) = 1; Gv(0, 2) = -v * dt * std::sin(swa + phi);
    Gv(1, 1) = 1; Gv(1, 2) = v * dt * std::cos(swa + phi);
    Gv(2, 2) = 1;
    Gu(0, 0) = dt * std::cos(swa + phi); Gu(0, 1) = -v * dt * std::sin(swa + phi);
    Gu(1, 0) = dt * std::sin(swa + phi); Gu(1, 1) = v * dt * std::cos(swa + phi);
    Gu(2, 0) = dt * std::sin(swa) / wb;  Gu(2, 1) = v * dt * std::cos(swa) / wb;

    Mat<T> Pvv(3, 3);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Pvv(i, j) = P(i, j);
    Mat<T> Pn = add(matmul(matmul(Gv, Pvv), transpose(Gv)), matmul(matmul(Gu, Q), transpose(Gu)));
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) P(i, j) = Pn(i, j);
    if (P.r > 3) {
        // Q2: literal block width is cols()-4, leaving the last column/row stale (EKF.cpp:442-443).
        const int width = (flags & FLAG_Q2_FULL_WIDTH) ? P.c - 3 : P.c - 4;
        for (int c = 3; c < 3 + width; c++) {
            T col[3] = {P(0, c), P(1, c), P(2, c)};
            for (int i = 0; i < 3; i++) {
                T s = 0;
                for (int k = 0; k < 3; k++) s += Gv(i, k) * col[k];
                P(i, c) = s;
            }
        }
        for (int c = 3; c < 3 + width; c++)
            for (int i = 0; i < 3; i++) P(c, i) = P(i, c);
    }
    X[0] = X[0] + v * dt * std::cos(swa + phi);
    X[1] = X[1] + v * dt * std::sin(swa + phi);
    X[2] = pi2pi(X[2] + v * dt * std::sin(swa) / wb);
}

// EKF.cpp:328-352 observeHeading.  dense=true runs slam.h:700-725 literally (O(n^3)).
template <class T>
void ekf_observe_heading(Vec<T>& X, Mat<T>& P, T phi, bool use_heading, bool dense) {
    if (!use_heading) return;
    // float sigmaPhi = 0.01F * pi / 180.0F evaluated in double then narrowed (EKF.cpp:337)
    const T sigma = static_cast<T>(0.01F * kPi / 180.0F);
    const T v = pi2pi(phi - X[2]);
    const T R = sigma * sigma;
    if (dense) {
        Mat<T> H(1, (int)X.size());
        H(0, 2) = T(1);
        Mat<T> Rm(1, 1);
        Rm(0, 0) = R;
        joseph_update_dense(X, P, Vec<T>{v}, Rm, H);
    } else {
        joseph_update_heading(X, P, v, R);
    }
}

// EKF.cpp:457-479 singleUpdate — re-linearised at the updated X for each observation.
// Z is 2 x m (column i = (range, bearing)), idf 1-based.  Returns #updates skipped.
template <class T>
int ekf_single_update(Vec<T>& X, Mat<T>& P, const Mat<T>& Z, const Mat<T>& R, const std::vector<int>& idf,
                      unsigned flags) {
    int skipped = 0;
    for (int i = 0; i < Z.c; i++) {
        ObserveModel<T> om = ekf_observe_model(X, idf[i]);
        Vec<T> V(2);
        V[0] = Z(0, i) - om.z[0];
        V[1] = pi2pi(Z(1, i) - om.z[1]);
        if (!cholesky_update(X, P, V, R, om.H, flags)) skipped++;
    }
    return skipped;
}

// EKF.cpp:93-129 batchUpdate — all observations linearised at the same X, one rank-2m update.
template <class T>
int ekf_batch_update(Vec<T>& X, Mat<T>& P, const Mat<T>& Z, const Mat<T>& R, const std::vector<int>& idf,
                     unsigned flags) {
    const int m = Z.c, n = (int)X.size();
    Mat<T> H(2 * m, n), RR(2 * m, 2 * m);
    Vec<T> V(2 * m);
    for (int i = 0; i < m; i++) {
        ObserveModel<T> om = ekf_observe_model(X, idf[i]);
        for (int j = 0; j < n; j++) {
            H(2 * i, j) = om.H(0, j);
            H(2 * i + 1, j) = om.H(1, j);
        }
        V[2 * i] = Z(0, i) - om.z[0];
        V[2 * i + 1] = pi2pi(Z(1, i) - om.z[1]);
        for (int a = 0; a < 2; a++)
            for (int b = 0; b < 2; b++) RR(2 * i + a, 2 * i + b) = R(a, b);
    }
    // with m == 0 every factor is empty and the update is a no-op (main.cpp:188 does call it so)
    if (m == 0) return 0;
    return cholesky_update(X, P, V, RR, H, flags) ? 0 : 1;
}

// EKF.cpp:481-496
template <class T>
int ekf_update(Vec<T>& X, Mat<T>& P, const Mat<T>& Z, const Mat<T>& R, const std::vector<int>& idf, bool batch,
               unsigned flags) {
    return batch ? ekf_batch_update(X, P, Z, R, idf, flags) : ekf_single_update(X, P, Z, R, idf, flags);
}

// EKF.cpp:28-91 addOneNewFeature — literal copy / resize / zero / copy-back growth.
template <class T>
void ekf_add_one_new_feature(Vec<T>& X, Mat<T>& P, T r, T b, const Mat<T>& R) {
    const int len = (int)X.size();
    const T s = std::sin(X[2] + b), c = std::cos(X[2] + b);
    X.push_back(X[0] + r * c);
    X.push_back(X[1] + r * s);
    Mat<T> Gv(2, 3), Gz(2, 2);
    Gv(0, 0) = 1; Gv(0, 2) = -r * s;
    Gv(1, 1) = 1; Gv(1, 2) = r * c;
    Gz(0, 0) = c; Gz(0, 1) = -r * s;
    Gz(1, 0) = s; Gz(1, 1) = r * c;
    Mat<T> AP = P;
    P = Mat<T>(len + 2, len + 2);
    for (int j = 0; j < len; j++)
        for (int i = 0; i < len; i++) P(i, j) = AP(i, j);
    Mat<T> Pvv(3, 3);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Pvv(i, j) = AP(i, j);
    Mat<T> Pff = add(matmul(matmul(Gv, Pvv), transpose(Gv)), matmul(matmul(Gz, R), transpose(Gz)));
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++) P(len + i, len + j) = Pff(i, j);
    Mat<T> GvP = matmul(Gv, Pvv);
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 3; j++) {
            P(len + i, j) = GvP(i, j);
            P(j, len + i) = GvP(i, j);
        }
    if (len > 3) {
        for (int j = 3; j < len; j++)
            for (int i = 0; i < 2; i++) {
                T acc = 0;
                for (int k = 0; k < 3; k++) acc += Gv(i, k) * P(k, j);
                P(len + i, j) = acc;
                P(j, len + i) = acc;
            }
    }
}

// EKF.cpp:9-26 augment
template <class T>
void ekf_augment(Vec<T>& X, Mat<T>& P, const Mat<T>& Z, const Mat<T>& R) {
    for (int i = 0; i < Z.c; i++) ekf_add_one_new_feature(X, P, Z(0, i), Z(1, i), R);
}

template <class T>
struct NormalizedInnovation {
    T nis, nd;
};

// EKF.cpp:131-144 computeAssociation — dense H*P*H^T (O(n^2) per pair), S not symmetrised.
template <class T>
NormalizedInnovation<T> ekf_compute_association(const Vec<T>& X, const Mat<T>& P, T zr, T zb, const Mat<T>& R,
                                               int idf) {
    ObserveModel<T> om = ekf_observe_model(X, idf);
    Mat<T> V(2, 1);
    V(0, 0) = zr - om.z[0];
    V(1, 0) = pi2pi(zb - om.z[1]);
    Mat<T> S = add(matmul(matmul(om.H, P), transpose(om.H)), R);
    Mat<T> nis = matmul(matmul(transpose(V), inverse(S)), V);
    const T nd = nis(0, 0) + std::log(determinant(S));
    return {nis(0, 0), nd};
}

// Same quantity using only the 5 x 5 sub-block of P that the sparse H touches
// (columns 0,1,2,f,f+1); products accumulated in the same column order as the dense
// form, so the two agree to rounding.  Used where the O(n^2)-per-pair form is infeasible.
template <class T>
NormalizedInnovation<T> ekf_compute_association_sparse(const Vec<T>& X, const Mat<T>& P, T zr, T zb,
                                                      const Mat<T>& R, int idf) {
    const int f = 3 + 2 * (idf - 1);
    const int cols[5] = {0, 1, 2, f, f + 1};
    ObserveModel<T> om;
    {
        const T dx = X[f] - X[0], dy = X[f + 1] - X[1];
        const T d2 = dx * dx + dy * dy, d = std::sqrt(d2);
        om.z[0] = d;
        om.z[1] = std::atan2(dy, dx) - X[2];
        om.H = Mat<T>(2, 5);
        om.H(0, 0) = -(dx / d);  om.H(0, 1) = -(dy / d);  om.H(0, 2) = T(0);
        om.H(1, 0) = dy / d2;    om.H(1, 1) = -(dx / d2); om.H(1, 2) = T(-1);
        om.H(0, 3) = dx / d;     om.H(0, 4) = dy / d;
        om.H(1, 3) = -(dy / d2); om.H(1, 4) = dx / d2;
    }
    Mat<T> Pc(5, 5);
    for (int a = 0; a < 5; a++)
        for (int b = 0; b < 5; b++) Pc(a, b) = P(cols[a], cols[b]);
    Mat<T> V(2, 1);
    V(0, 0) = zr - om.z[0];
    V(1, 0) = pi2pi(zb - om.z[1]);
    Mat<T> S = add(matmul(matmul(om.H, Pc), transpose(om.H)), R);
    Mat<T> nis = matmul(matmul(transpose(V), inverse(S)), V);
    const T nd = nis(0, 0) + std::log(determinant(S));
    return {nis(0, 0), nd};
}

template <class T>
struct Association {
    Mat<T> ZF, ZN;
    std::vector<int> idf;
    // per-observation decisions (the GPU contract): jbest (0 = none), outer = min_j nis, is_new
    std::vector<int> jbest;
    std::vector<T> nbest, outer;
    std::vector<uint8_t> is_new;
};

// EKF.cpp:235-326 dataAssociate.  dense=false swaps computeAssociation for its sparse twin.
template <class T>
Association<T> ekf_data_associate(const Vec<T>& X, const Mat<T>& P, const Mat<T>& Z, const Mat<T>& R, T gate1,
                                  T gate2, unsigned flags, bool dense) {
    Association<T> out;
    const int nf = ((int)X.size() - 3) / 2;
    std::vector<int> zf_cols, zn_cols;
    for (int i = 0; i < Z.c; i++) {
        int jbest = 0;
        T nbest = std::numeric_limits<T>::infinity();
        T outer = std::numeric_limits<T>::infinity();
        for (int j = 1; j <= nf; j++) {
            NormalizedInnovation<T> ni = dense ? ekf_compute_association(X, P, Z(0, i), Z(1, i), R, j)
                                               : ekf_compute_association_sparse(X, P, Z(0, i), Z(1, i), R, j);
            if (ni.nis < gate1 && ni.nd < nbest) {  // EKF.cpp:275-283 (Q4)
                nbest = ni.nd;
                jbest = j;
            } else if (ni.nis < outer) {
                outer = ni.nis;
            }
        }
        bool is_new = false;
        if (jbest != 0) {
            zf_cols.push_back(i);
            out.idf.push_back(jbest);
        } else if (outer > gate2) {
            zn_cols.push_back(i);
            is_new = true;
        }
        out.jbest.push_back(jbest);
        out.nbest.push_back(nbest);
        out.outer.push_back(outer);
        out.is_new.push_back(is_new ? 1 : 0);
    }
    out.ZF = Mat<T>(2, (int)zf_cols.size());
    for (size_t k = 0; k < zf_cols.size(); k++) {
        out.ZF(0, (int)k) = Z(0, zf_cols[k]);
        out.ZF(1, (int)k) = Z(1, zf_cols[k]);
    }
    // Q5: the reference fills a shadowed local and returns the outer, EMPTY ZN (EKF.cpp:308-315,325)
    if (flags & FLAG_Q5_RETURN_ZN) {
        out.ZN = Mat<T>(2, (int)zn_cols.size());
        for (size_t k = 0; k < zn_cols.size(); k++) {
            out.ZN(0, (int)k) = Z(0, zn_cols[k]);
            out.ZN(1, (int)k) = Z(1, zn_cols[k]);
        }
    } else {
        out.ZN = Mat<T>(0, 0);
    }
    return out;
}

// EKF.cpp:146-233 dataAssociateTable (known associations; table holds 1-based map slots, 0 = unseen)
template <class T>
Association<T> ekf_data_associate_table(const Vec<T>& X, const Mat<T>& Z, const std::vector<int>& idz,
                                        std::vector<int>& table) {
    Association<T> out;
    std::vector<int> zf_cols, zn_cols, idn;
    for (size_t i = 0; i < idz.size(); i++) {
        const int id = idz[i];
        if (table[id - 1] == 0) {
            zn_cols.push_back((int)i);
            idn.push_back(id);
        } else {
            zf_cols.push_back((int)i);
            out.idf.push_back(table[id - 1]);
        }
    }
    out.ZF = Mat<T>(zf_cols.empty() ? 0 : 2, (int)zf_cols.size());
    for (size_t k = 0; k < zf_cols.size(); k++) {
        out.ZF(0, (int)k) = Z(0, zf_cols[k]);
        out.ZF(1, (int)k) = Z(1, zf_cols[k]);
    }
    out.ZN = Mat<T>(zn_cols.empty() ? 0 : 2, (int)zn_cols.size());
    for (size_t k = 0; k < zn_cols.size(); k++) {
        out.ZN(0, (int)k) = Z(0, zn_cols[k]);
        out.ZN(1, (int)k) = Z(1, zn_cols[k]);
    }
    const int nf = (int)(((int)X.size() - 3) / 2.0F);
    for (size_t i = 0; i < idn.size(); i++) table[idn[i] - 1] = nf + (int)i + 1;
    return out;
}

// ---------------------------------------------------------------------------------
// Simulator helpers (slam.h) — needed to reproduce config 1's input stream.
// ---------------------------------------------------------------------------------
template <class T>
int signum(T x) {  // slam.h:924-928
    return (T(0) < x) - (x < T(0));
}

// slam.h:279-332 computeSWA
template <class T>
void compute_swa(const Vec<T>& X, const Mat<T>& WP, int& iwp, T minD, T& swa, T rateSWA, T maxSWA, T dt) {
    if (WP.c <= 0) return;
    T cx = WP(0, iwp - 1), cy = WP(1, iwp - 1);
    const T d2 = (cx - X[0]) * (cx - X[0]) + (cy - X[1]) * (cy - X[1]);
    if (d2 < minD * minD) {
        iwp = iwp + 1;
        if (iwp > WP.c) {
            iwp = 0;
            return;
        }
        cx = WP(0, iwp - 1);
        cy = WP(1, iwp - 1);
    }
    T deltaG = pi2pi(std::atan2(cy - X[1], cx - X[0]) - X[2] - swa);
    const T maxDelta = rateSWA * dt;
    // Q18 (simulator only): the reference calls signum<int>(float) — the argument is TRUNCATED
    // to int first (slam.h:317,324), so any |value| < 1 has signum 0.
    if (std::abs(deltaG) > maxDelta) deltaG = maxDelta * signum<int>(static_cast<int>(deltaG));
    swa = swa + deltaG;
    if (std::abs(swa) > maxSWA) swa = signum<int>(static_cast<int>(swa)) * maxSWA;
}

// slam.h:952-966 vehicleModel
template <class T>
void vehicle_model(Vec<T>& X, T v, T swa, T wb, T dt) {
    const T x0 = X[0], x1 = X[1], x2 = X[2];
    X[0] = x0 + v * dt * std::cos(swa + x2);
    X[1] = x1 + v * dt * std::sin(swa + x2);
    X[2] = pi2pi(x2 + v * dt * std::sin(swa) / wb);
}

// slam.h:575-582 getObservations -> :608-683 getVisibleLandmarks -> :339-368 computeRangeBearing.
// The visibility test is done in double in the reference regardless of T (slam.h:623-648).
template <class T>
void get_observations(const Vec<T>& X, const Mat<T>& LM, const std::vector<int>& tags, T rmax, Mat<T>& Z,
                      std::vector<int>& idf) {
    std::vector<int> vis;
    const double phi = X[2];
    for (int i = 0; i < LM.c; i++) {
        const double dx = LM(0, i) - X[0], dy = LM(1, i) - X[1];
        const double rm = static_cast<double>(rmax);
        if ((std::abs(dx) < rm && std::abs(dy) < rm) && ((dx * std::cos(phi) + dy * std::sin(phi)) > 0.0) &&
            ((dx * dx + dy * dy) < rm * rm))
            vis.push_back(i);
    }
    Z = Mat<T>(vis.empty() ? 0 : 2, (int)vis.size());
    idf.clear();
    for (size_t k = 0; k < vis.size(); k++) {
        const T dx = LM(0, vis[k]) - X[0], dy = LM(1, vis[k]) - X[1];
        Z(0, (int)k) = std::sqrt(dx * dx + dy * dy);
        Z(1, (int)k) = std::atan2(dy, dx) - X[2];
        idf.push_back(tags[vis[k]]);
    }
}

// ---------------------------------------------------------------------------------
// Particle filter (slam/src/PF.cpp)
// ---------------------------------------------------------------------------------
template <class T>
struct Particle {  // slam.h:120-127
    T w = 0;
    Vec<T> X = Vec<T>(3, T(0));
    Mat<T> P = Mat<T>(3, 3);
    Mat<T> XF;               // 2 x Nf
    std::vector<Mat<T>> PF;  // Nf of 2 x 2
};

// PF.cpp:319-341
template <class T>
std::vector<Particle<T>> pf_initialize_particles(int n) {
    std::vector<Particle<T>> ps(n);
    for (auto& p : ps) {
        p.w = T(1) / static_cast<T>(n);
        p.XF = Mat<T>(0, 0);
    }
    return ps;
}

// PF.cpp:419-471
template <class T>
void pf_predict(Particle<T>& p, T v, T swa, const Mat<T>& Q, T wb, T dt) {
    const T phi = p.X[2];
    Mat<T> Gv(3, 3), Gu(3, 2);
    Gv(0, 0) = 1; Gv(0, 2) = -v * dt * std::sin(swa + phi);
    Gv(1, 1) = 1; Gv(1, 2) = v * dt * std::cos(swa + phi);
    Gv(2, 2) = 1;
    Gu(0, 0) = dt * std::cos(swa + phi); Gu(0, 1) = -v * dt * std::sin(swa + phi);
    Gu(1, 0) = dt * std::sin(swa + phi); Gu(1, 1) = v * dt * std::cos(swa + phi);
    Gu(2, 0) = dt * std::sin(swa) / wb;  Gu(2, 1) = v * dt * std::cos(swa) / wb;
    p.P = add(matmul(matmul(Gv, p.P), transpose(Gv)), matmul(matmul(Gu, Q), transpose(Gu)));
    p.X[0] = p.X[0] + v * dt * std::cos(swa + phi);
    p.X[1] = p.X[1] + v * dt * std::sin(swa + phi);
    p.X[2] = pi2pi(p.X[2] + v * dt * std::sin(swa) / wb);
}

// PF.cpp:382-417
template <class T>
void pf_observe_heading(Particle<T>& p, T phi, bool use_heading) {
    if (!use_heading) return;
    const T sigma = static_cast<T>(0.01F * kPi / 180.0F);
    Mat<T> H(1, 3), R(1, 1);
    H(0, 2) = T(1);
    R(0, 0) = sigma * sigma;
    joseph_update_dense(p.X, p.P, Vec<T>{pi2pi(phi - p.X[2])}, R, H);
}

template <class T>
struct Jacobians {
    T zp[2];
    Mat<T> Hv, Hf, Sf;
};

// PF.cpp:70-135 computeJacobians for ONE feature id (1-based); bearing IS wrapped here.
template <class T>
Jacobians<T> pf_compute_jacobians(const Particle<T>& p, int idf, const Mat<T>& R) {
    const int id = idf - 1;
    Jacobians<T> J;
    const T dx = p.XF(0, id) - p.X[0], dy = p.XF(1, id) - p.X[1];
    const T d2 = dx * dx + dy * dy, d = std::sqrt(d2);
    J.zp[0] = d;
    J.zp[1] = pi2pi(std::atan2(dy, dx) - p.X[2]);
    J.Hv = Mat<T>(2, 3);
    J.Hv(0, 0) = -dx / d; J.Hv(0, 1) = -dy / d; J.Hv(0, 2) = T(0);
    J.Hv(1, 0) = dy / d2; J.Hv(1, 1) = -dx / d2; J.Hv(1, 2) = T(-1);
    J.Hf = Mat<T>(2, 2);
    J.Hf(0, 0) = dx / d;   J.Hf(0, 1) = dy / d;
    J.Hf(1, 0) = -dy / d2; J.Hf(1, 1) = dx / d2;
    J.Sf = add(matmul(matmul(J.Hf, p.PF[id]), transpose(J.Hf)), R);
    return J;
}

// PF.cpp:279-317 gaussEvaluate (logFlag=false branch; the log branch has no callers)
template <class T>
T pf_gauss_evaluate(const Vec<T>& V, const Mat<T>& S, unsigned flags) {
    const int D = (int)V.size();
    Mat<T> L = cholesky_decomposition(S);
    Mat<T> SC = transpose(L);  // upper U = L^T (PF.cpp:287-289)
    // Q9: literal nin = U^-1 V (metric L^T L); intended nin = L^-1 V (metric S)
    Mat<T> inv = (flags & FLAG_Q9_METRIC_S) ? inverse(L) : inverse(SC);
    T sum = 0;
    for (int i = 0; i < D; i++) {
        T s = 0;
        for (int k = 0; k < D; k++) s += inv(i, k) * V[k];
        sum += s * s;
    }
    const T E = T(-0.5) * sum;
    T prod = 1;
    for (int i = 0; i < D; i++) prod *= SC(i, i);
    const T C = std::pow(T(2) * static_cast<T>(kPi), static_cast<T>(D) / T(2)) * prod;
    return std::exp(E) / C;
}

// PF.cpp:343-359
template <class T>
T pf_likelihood(const Particle<T>& p, const Mat<T>& Z, const std::vector<int>& idf, const Mat<T>& R,
                unsigned flags) {
    T w = 1;
    for (size_t i = 0; i < idf.size(); i++) {
        Jacobians<T> J = pf_compute_jacobians(p, idf[i], R);
        Vec<T> V{Z(0, (int)i) - J.zp[0], pi2pi(Z(1, (int)i) - J.zp[1])};
        w = w * pf_gauss_evaluate(V, J.Sf, flags);
    }
    return w;
}

// PF.cpp:502-544 sampleProposal.  xi = the three standard-normal draws that
// multivariateNormalGaussianDistribution (slam.h:753-764) would take from Boost (Q7: an INPUT here).
template <class T>
void pf_sample_proposal(Particle<T>& p, const Mat<T>& Z, const std::vector<int>& idf, const Mat<T>& R,
                        const T xi[3], unsigned flags) {
    Vec<T> X = p.X;
    Mat<T> P = p.P;
    const Vec<T> X0 = X;
    const Mat<T> P0 = P;
    for (size_t k = 0; k < idf.size(); k++) {
        Jacobians<T> J = pf_compute_jacobians(p, idf[k], R);
        Mat<T> Sfi = inverse(J.Sf);
        Mat<T> V(2, 1);
        V(0, 0) = Z(0, (int)k) - J.zp[0];
        V(1, 0) = pi2pi(Z(1, (int)k) - J.zp[1]);
        Mat<T> HvT = transpose(J.Hv);
        Mat<T> PT = add(matmul(matmul(HvT, Sfi), J.Hv), inverse(P));
        P = inverse(PT);
        Mat<T> dX = matmul(matmul(matmul(P, HvT), Sfi), V);
        for (int i = 0; i < 3; i++) X[i] = X[i] + dX(i, 0);
        p.X = X;
        p.P = P;
    }
    // slam.h:753-764: XS = chol(P) * xi + X
    Mat<T> Lp = cholesky_decomposition(P);
    Vec<T> XS(3);
    for (int i = 0; i < 3; i++) {
        T s = 0;
        for (int k = 0; k < 3; k++) s += Lp(i, k) * xi[k];
        XS[i] = s + X[i];
    }
    p.X = XS;
    p.P = Mat<T>(3, 3);
    const T like = pf_likelihood(p, Z, idf, R, flags);
    Vec<T> d0{X0[0] - XS[0], X0[1] - XS[1], pi2pi(X0[2] - XS[2])};  // PF.cpp:62-68 computeDelta
    Vec<T> d1{X[0] - XS[0], X[1] - XS[1], pi2pi(X[2] - XS[2])};
    const T prior = pf_gauss_evaluate(d0, P0, flags);
    const T prop = pf_gauss_evaluate(d1, P, flags);
    p.w = p.w * like * prior / prop;
}

// PF.cpp:222-277 featureUpdate — Jacobians at the sampled pose, each feature independently.
template <class T>
void pf_feature_update(Particle<T>& p, const Mat<T>& Z, const std::vector<int>& idf, const Mat<T>& R,
                       unsigned flags) {
    const int len = (int)idf.size();
    std::vector<Vec<T>> XF(len);
    std::vector<Mat<T>> PF(len);
    std::vector<Jacobians<T>> J(len);
    for (int i = 0; i < len; i++) {
        const int id = idf[i] - 1;
        XF[i] = Vec<T>{p.XF(0, id), p.XF(1, id)};
        PF[i] = p.PF[id];
        J[i] = pf_compute_jacobians(p, idf[i], R);
    }
    for (int i = 0; i < len; i++) {
        Vec<T> V{Z(0, i) - J[i].zp[0], pi2pi(Z(1, i) - J[i].zp[1])};
        cholesky_update(XF[i], PF[i], V, R, J[i].Hf, flags);
    }
    for (int i = 0; i < len; i++) {
        const int id = idf[i] - 1;
        p.XF(0, id) = XF[i][0];
        p.XF(1, id) = XF[i][1];
        p.PF[id] = PF[i];
    }
}

// PF.cpp:9-60 addOneNewFeature (particle overload): pose treated as exact.
template <class T>
void pf_add_new_features(Particle<T>& p, const Mat<T>& Z, const Mat<T>& R) {
    const int m = Z.c;
    if (m <= 0) return;
    const int old = p.XF.c;
    Mat<T> XF(2, old + m);
    for (int j = 0; j < old; j++) {
        XF(0, j) = p.XF(0, j);
        XF(1, j) = p.XF(1, j);
    }
    for (int i = 0; i < m; i++) {
        const T r = Z(0, i), b = Z(1, i);
        const T s = std::sin(p.X[2] + b), c = std::cos(p.X[2] + b);
        XF(0, old + i) = p.X[0] + r * c;
        XF(1, old + i) = p.X[1] + r * s;
        Mat<T> Gz(2, 2);
        Gz(0, 0) = c; Gz(0, 1) = -r * s;
        Gz(1, 0) = s; Gz(1, 1) = r * c;
        p.PF.push_back(matmul(matmul(Gz, R), transpose(Gz)));
    }
    p.XF = XF;
}

template <class T>
struct Stratified {
    std::vector<int> keep;  // 0-based (Q11)
    T neff;
};

// Canonical summation order of the INTENDED resampler ("radix-32 hierarchical Kogge-Stone"),
// shared bit-for-bit with conan_slam_b200/csrc/pf.cu so that cumulative weights — hence
// resampled indices — are identical on CPU and on 1/2/4/8 GPUs:
//   level 0: each aligned group of 32 consecutive values gets an inclusive Kogge-Stone scan
//            (x_i <- x_i + x_{i-d} for d = 1,2,4,8,16, all lanes stepping together);
//   level l: the group totals of level l-1 are scanned the same way;
//   cum(i)  = ((0 + o_top) + ... + o_1) + scan_0[i], o_l = exclusive offset of i's group at level l.
template <class T>
void ks32_group_scan(Vec<T>& x) {  // in place, every aligned group of 32
    const size_t n = x.size();
    for (size_t g0 = 0; g0 < n; g0 += 32) {
        T v[32];
        for (int i = 0; i < 32; i++) v[i] = (g0 + i < n) ? x[g0 + i] : T(0);
        for (int d = 1; d < 32; d <<= 1) {
            T y[32];
            for (int i = 0; i < 32; i++) y[i] = v[i];
            for (int i = d; i < 32; i++) v[i] = y[i] + y[i - d];
        }
        for (int i = 0; i < 32 && g0 + i < n; i++) x[g0 + i] = v[i];
    }
}
template <class T>
std::vector<Vec<T>> canonical_scan_levels(const Vec<T>& in) {
    std::vector<Vec<T>> lv;
    lv.push_back(in);
    while (true) {
        Vec<T>& cur = lv.back();
        ks32_group_scan(cur);
        if (cur.size() <= 32) break;
        Vec<T> tot((cur.size() + 31) / 32);
        for (size_t g = 0; g < tot.size(); g++) tot[g] = cur[std::min(cur.size() - 1, g * 32 + 31)];
        lv.push_back(tot);
    }
    return lv;
}
template <class T>
T canonical_sum(const Vec<T>& in) {
    std::vector<Vec<T>> lv = canonical_scan_levels(in);
    return lv.back().back();
}
template <class T>
Vec<T> canonical_cumsum(const Vec<T>& in) {
    std::vector<Vec<T>> lv = canonical_scan_levels(in);
    Vec<T> out(in.size());
    for (size_t i = 0; i < in.size(); i++) {
        T off = 0;
        for (int l = (int)lv.size() - 1; l >= 1; l--) {
            const size_t gi = i >> (5 * l);
            if (gi & 31) off = off + lv[l][gi - 1];
        }
        out[i] = off + lv[0][i];
    }
    return out;
}

// PF.cpp:546-577 stratifiedResample with PF.cpp:579-596 stratifiedRandom folded in.
// u[i] = the per-slot random deviate (Q12: an INPUT; the reference draws N(0,1)); the
// comb k/2 + i*k is accumulated by repeated addition exactly as the reference does.
// W is normalised and overwritten by its running sum, like the reference's by-ref W.
// REF_LITERAL: plain left-to-right sums and the literal slot loop (Q10).
// FLAG_Q10_SEARCH: per-slot first-true search, sums in the canonical order above.
template <class T>
Stratified<T> pf_stratified_resample(Vec<T>& W, const Vec<T>& u, unsigned flags) {
    const int len = (int)W.size();
    const bool intended = (flags & FLAG_Q10_SEARCH) != 0;
    T ws = 0;
    if (intended) ws = canonical_sum(W);
    else for (int i = 0; i < len; i++) ws += W[i];
    for (int i = 0; i < len; i++) W[i] = W[i] / ws;
    T s2 = 0;
    if (intended) {
        Vec<T> sq(len);
        for (int i = 0; i < len; i++) sq[i] = W[i] * W[i];
        s2 = canonical_sum(sq);
    } else {
        for (int i = 0; i < len; i++) s2 += W[i] * W[i];
    }
    Stratified<T> out;
    out.neff = T(1) / s2;
    out.keep.assign(len, 0);
    const T k = T(1) / static_cast<T>(len);
    Vec<T> select(len);
    T di = k / T(2);
    for (int i = 0; i < len; i++) {
        if (i > 0) di = di + k;
        select[i] = di + u[i] * (k - k / T(2));
    }
    if (intended) {
        W = canonical_cumsum(W);
        // Keep[c] = min{ i : select[c] < cumW[i] }  (cumW is nondecreasing -> first-true search)
        for (int c = 0; c < len; c++) {
            int lo = 0, hi = len;
            while (lo < hi) {
                const int mid = (lo + hi) / 2;
                if (select[c] < W[mid]) hi = mid; else lo = mid + 1;
            }
            out.keep[c] = lo < len ? lo : len - 1;
        }
    } else {
        T summer = W[0];  // PF.cpp:559-564
        for (int i = 1; i < len; i++) {
            summer = summer + W[i];
            W[i] = summer;
        }
        int ctr = 1;  // literal PF.cpp:566-574 (Q10): condition ignores ctr -> first hit takes all slots
        for (int i = 0; i < len; i++)
            while (ctr <= len && select[i] < W[i]) {
                out.keep[ctr - 1] = i;
                ctr++;
            }
    }
    return out;
}

// PF.cpp:473-500 resampleParticles.  Parity is pinned at the stratifiedResample boundary
// (Q11: the literal copy indexes particles[keep-1], UB for keep==0); the copy here uses
// keep[] directly (0-based), i.e. the INTENDED gather.  Returns whether a resample happened.
template <class T>
bool pf_resample_particles(std::vector<Particle<T>>& ps, const Vec<T>& u, T num_effective, bool resample_on,
                           unsigned flags, Stratified<T>* strat_out = nullptr) {
    const int n = (int)ps.size();
    Vec<T> W(n);
    for (int i = 0; i < n; i++) W[i] = ps[i].w;
    T ws = 0;
    if (flags & FLAG_Q10_SEARCH) ws = canonical_sum(W);
    else for (int i = 0; i < n; i++) ws += W[i];
    for (int i = 0; i < n; i++) { W[i] = W[i] / ws; ps[i].w = ps[i].w / ws; }
    Stratified<T> st = pf_stratified_resample(W, u, flags);
    if (strat_out) *strat_out = st;
    if (st.neff < num_effective && resample_on) {
        std::vector<Particle<T>> np;
        np.reserve(n);
        for (int i = 0; i < n; i++) {
            Particle<T> q = ps[st.keep[i]];
            q.w = T(1) / static_cast<T>(n);
            np.push_back(q);
        }
        ps = np;
        return true;
    }
    return false;
}

}  // namespace oracle
