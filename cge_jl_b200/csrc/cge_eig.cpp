// cge_eig.cpp -- host-only: principal axis of a symmetric d x d matrix, the one O(d^3) step of a cluster cut
// in the landmark selection (cge_select.cu; /root/reference/src/landmarks.jl:160-162 hands it to LAPACK:
// eigvecs(y'wy)[:, end]).  Used when the caller passes no LAPACK callback to cge_b200_landmarks_select.
//
// Householder reduction to tridiagonal form working on the LOWER triangle only (as LAPACK's dsytd2: the
// symmetric matrix-vector product and the rank-2 update touch i >= j, 2/3 d^3 multiply-adds), bisection on
// the Sturm sequence for the largest eigenvalue, inverse iteration on the tridiagonal matrix (LU with
// partial pivoting), back-transformation through the reflectors.  The inner loops run along rows
// (contiguous), compiled twice -- baseline x86-64 and AVX2+FMA -- and dispatched by the CPU at load time.
#include <algorithm>
#include <cmath>
#include <vector>

#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define CGE_CLONES __attribute__((target_clones("avx2,fma", "default")))
#else
#define CGE_CLONES
#endif

namespace cge {

namespace {

// A (row-major, leading dimension d) holds the lower triangle; on return its diagonal / subdiagonal are
// those of the tridiagonal matrix T = Q^T A Q, column k below the subdiagonal holds the reflector
// u_k (u_k[k+1] = 1 implied), Q = H_0 H_1 ... H_{d-3}, H_k = I - tau_k u_k u_k^T.
CGE_CLONES void tridiagonalize_lower(double *A, int d, double *tau, double *u, double *p) {
    for (int k = 0; k + 2 < d; ++k) {
        double sigma = 0.0;
        for (int i = k + 2; i < d; ++i) sigma += A[(size_t)i * d + k] * A[(size_t)i * d + k];
        const double alpha = A[(size_t)(k + 1) * d + k];
        tau[k] = 0.0;
        if (sigma == 0.0) continue;
        const double nrm = std::sqrt(alpha * alpha + sigma);
        const double beta = alpha > 0.0 ? -nrm : nrm;
        const double u0 = alpha - beta;
        tau[k] = (beta - alpha) / beta;
        const double inv = 1.0 / u0;
        for (int i = k + 2; i < d; ++i) A[(size_t)i * d + k] *= inv;
        A[(size_t)(k + 1) * d + k] = beta;
        const int m = d - k - 1;  // trailing block B = A[k+1.., k+1..], lower triangle
        u[0] = 1.0;
        for (int i = 1; i < m; ++i) u[i] = A[(size_t)(k + 1 + i) * d + k];
        // p = tau * B u with B symmetric, lower triangle stored: row i gives sum_{j<=i} B_ij u_j to p_i and
        // u_i B_ij to p_j (j < i)
        for (int i = 0; i < m; ++i) p[i] = 0.0;
        for (int i = 0; i < m; ++i) {
            const double *row = A + (size_t)(k + 1 + i) * d + (k + 1);
            const double ui = u[i];
            double s = 0.0;
            for (int j = 0; j < i; ++j) {
                s += row[j] * u[j];
                p[j] += ui * row[j];
            }
            p[i] += s + row[i] * ui;
        }
        double up = 0.0;
        for (int i = 0; i < m; ++i) {
            p[i] *= tau[k];
            up += u[i] * p[i];
        }
        const double kk = 0.5 * tau[k] * up;
        for (int i = 0; i < m; ++i) p[i] -= kk * u[i];  // q = p - (tau/2)(u.p) u
        for (int i = 0; i < m; ++i) {                   // B <- B - u q^T - q u^T, lower triangle
            double *row = A + (size_t)(k + 1 + i) * d + (k + 1);
            const double ui = u[i], qi = p[i];
            for (int j = 0; j <= i; ++j) row[j] -= ui * p[j] + qi * u[j];
        }
    }
}

}  // namespace

// Eigenvector of the LARGEST eigenvalue of the symmetric matrix a (row-major d x d, upper triangle read),
// unit length, largest-magnitude component positive; *lambda_out (optional) receives the eigenvalue.
void sym_top_eigvec(const double *a_in, int d, double *v_out, double *lambda_out) {
    std::vector<double> A((size_t)d * d);
    for (int i = 0; i < d; ++i)
        for (int j = i; j < d; ++j) A[(size_t)j * d + i] = a_in[(size_t)i * d + j];  // lower <- upper
    std::vector<double> diag(d), off(std::max(d, 1), 0.0), tau(std::max(d, 1), 0.0), u(d), p(d);
    tridiagonalize_lower(A.data(), d, tau.data(), u.data(), p.data());
    for (int i = 0; i < d; ++i) diag[i] = A[(size_t)i * d + i];
    for (int i = 0; i + 1 < d; ++i) off[i] = A[(size_t)(i + 1) * d + i];
    // largest eigenvalue of T by bisection (Sturm count of eigenvalues < x)
    double lo = diag[0], hi = diag[0];
    for (int i = 0; i < d; ++i) {
        const double r = (i > 0 ? std::fabs(off[i - 1]) : 0.0) + (i + 1 < d ? std::fabs(off[i]) : 0.0);
        lo = std::min(lo, diag[i] - r);
        hi = std::max(hi, diag[i] + r);
    }
    const double scale = std::max(std::fabs(lo), std::fabs(hi));
    const double tiny = std::max(scale, 1e-300) * 1e-300 + 1e-300;
    auto count_below = [&](double xv) {
        int c = 0;
        double q = diag[0] - xv;
        if (q < 0.0) ++c;
        for (int i = 1; i < d; ++i) {
            if (q == 0.0) q = tiny;
            q = diag[i] - xv - off[i - 1] * off[i - 1] / q;
            if (q < 0.0) ++c;
        }
        return c;
    };
    for (int it = 0; it < 200; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (mid <= lo || mid >= hi) break;
        if (count_below(mid) >= d) hi = mid; else lo = mid;  // all d eigenvalues below mid -> go down
    }
    const double lam = 0.5 * (lo + hi);
    // inverse iteration on T - lam I (tridiagonal LU with partial pivoting, factored once)
    std::vector<double> y(d), a1(d), b1(d), c1(d), c2(d);
    std::vector<char> swp(d, 0);
    const double eps = 2.220446049250313e-16;
    const double pert = std::max(scale, 1e-300) * eps;
    for (int i = 0; i < d; ++i) y[i] = 1.0 + 0.01 * ((i * 7919) % 13);
    for (int i = 0; i < d; ++i) {
        b1[i] = diag[i] - lam;
        c1[i] = i + 1 < d ? off[i] : 0.0;
        c2[i] = 0.0;
    }
    for (int i = 0; i + 1 < d; ++i) {
        const double sub = off[i];
        if (std::fabs(sub) > std::fabs(b1[i])) {  // swap rows i, i+1
            swp[i] = 1;
            const double nb = sub, nc = b1[i + 1], nc2 = c1[i + 1];
            const double ob = b1[i], oc = c1[i];
            b1[i] = nb; c1[i] = nc; c2[i] = nc2;
            const double l = ob / nb;
            a1[i] = l;
            b1[i + 1] = oc - l * nc;
            c1[i + 1] = -l * nc2;
        } else {
            if (b1[i] == 0.0) b1[i] = pert;
            const double l = sub / b1[i];
            a1[i] = l;
            b1[i + 1] -= l * c1[i];
        }
    }
    if (b1[d - 1] == 0.0) b1[d - 1] = pert;
    for (int iter = 0; iter < 4; ++iter) {
        for (int i = 0; i + 1 < d; ++i) {  // forward
            if (swp[i]) std::swap(y[i], y[i + 1]);
            y[i + 1] -= a1[i] * y[i];
        }
        for (int i = d - 1; i >= 0; --i) {  // backward
            double s = y[i];
            if (i + 1 < d) s -= c1[i] * y[i + 1];
            if (i + 2 < d) s -= c2[i] * y[i + 2];
            y[i] = s / b1[i];
        }
        double nrm = 0.0;
        for (int i = 0; i < d; ++i) nrm = std::max(nrm, std::fabs(y[i]));
        if (!(nrm > 0.0) || !std::isfinite(nrm)) {
            for (int i = 0; i < d; ++i) y[i] = i == 0 ? 1.0 : 0.0;
            break;
        }
        for (int i = 0; i < d; ++i) y[i] /= nrm;
    }
    // back-transform: v = H_0 H_1 ... H_{d-3} y
    for (int k = d - 3; k >= 0; --k) {
        if (tau[k] == 0.0) continue;
        double s = y[k + 1];
        for (int i = k + 2; i < d; ++i) s += A[(size_t)i * d + k] * y[i];
        s *= tau[k];
        y[k + 1] -= s;
        for (int i = k + 2; i < d; ++i) y[i] -= s * A[(size_t)i * d + k];
    }
    double n2 = 0.0;
    for (int i = 0; i < d; ++i) n2 += y[i] * y[i];
    n2 = std::sqrt(n2);
    int big = 0;
    for (int i = 1; i < d; ++i)
        if (std::fabs(y[i]) > std::fabs(y[big])) big = i;
    const double sgn = y[big] < 0.0 ? -1.0 : 1.0;
    for (int i = 0; i < d; ++i) v_out[i] = sgn * y[i] / n2;
    if (lambda_out) *lambda_out = lam;
}

}  // namespace cge
