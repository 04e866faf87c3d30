// tests/host_unit.cpp -- TEST INFRASTRUCTURE: unit checks of the host-side algebra of daliti_b200/csrc/host against
// straightforward restatements, bit for bit.  Built and run by tests/test_host_unit.py against the emulator build
// of the library (the host code is the same in both builds).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "eskf_lio_host.hpp"

using namespace dlt_host;

// the textbook routines the fixed-size / AVX2-cloned instances must reproduce exactly
static bool ref_invert(const double *A, int n, double *out) {
    std::vector<double> w((size_t)n * 2 * n);
    const int W = 2 * n;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            w[i * W + j] = A[i * n + j];
            w[i * W + n + j] = (i == j) ? 1.0 : 0.0;
        }
    for (int c = 0; c < n; c++) {
        int p = c;
        for (int r = c + 1; r < n; r++)
            if (std::fabs(w[r * W + c]) > std::fabs(w[p * W + c])) p = r;
        if (w[p * W + c] == 0.0) return false;
        if (p != c)
            for (int j = 0; j < W; j++) std::swap(w[p * W + j], w[c * W + j]);
        const double inv = 1.0 / w[c * W + c];
        for (int j = 0; j < W; j++) w[c * W + j] *= inv;
        for (int r = 0; r < n; r++) {
            if (r == c) continue;
            const double f = w[r * W + c];
            if (f == 0.0) continue;
            for (int j = 0; j < W; j++) w[r * W + j] -= f * w[c * W + j];
        }
    }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) out[i * n + j] = w[i * W + n + j];
    return true;
}
static bool ref_solve_first_columns(const double *A, int n, int m, double *X) {
    const int W = n + m;
    std::vector<double> w((size_t)n * W);
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < n; j++) w[i * W + j] = A[i * n + j];
        for (int j = 0; j < m; j++) w[i * W + n + j] = (i == j) ? 1.0 : 0.0;
    }
    for (int c = 0; c < n; c++) {
        int p = c;
        for (int r = c + 1; r < n; r++)
            if (std::fabs(w[r * W + c]) > std::fabs(w[p * W + c])) p = r;
        if (w[p * W + c] == 0.0) return false;
        if (p != c)
            for (int j = c; j < W; j++) std::swap(w[p * W + j], w[c * W + j]);
        const double piv = w[c * W + c];
        for (int r = c + 1; r < n; r++) {
            const double f = w[r * W + c] / piv;
            if (f == 0.0) continue;
            for (int j = c + 1; j < W; j++) w[r * W + j] -= f * w[c * W + j];
        }
    }
    for (int c = n - 1; c >= 0; c--) {
        const double piv = w[c * W + c];
        for (int j = 0; j < m; j++) {
            double s = w[c * W + n + j];
            for (int k = c + 1; k < n; k++) s -= w[c * W + k] * X[k * m + j];
            X[c * m + j] = s / piv;
        }
    }
    return true;
}

int main() {
    int failures = 0;
    std::mt19937_64 g(20261017);
    std::normal_distribution<double> N(0, 1);
    // ---- 24 x 24 (the fixed-size instances) and 10 x 10 (the generic path)
    for (int n : {24, 10}) {
        const int m = n / 2;
        std::vector<double> B(n * n), A(n * n), X(n * n), Y(n * n);
        for (int t = 0; t < 200; t++) {
            for (auto &v : B) v = N(g);
            for (int i = 0; i < n; i++)
                for (int j = 0; j < n; j++) {
                    double s = (i == j) ? 0.05 : 0.0;
                    for (int k = 0; k < n; k++) s += B[i * n + k] * B[j * n + k];
                    A[i * n + j] = s;
                }
            if (t % 3 == 0)  // the structure (P/R)^-1 has early on: blocks that do not couple
                for (int i = m; i < n; i++)
                    for (int j = 0; j < n; j++)
                        if (i != j) A[i * n + j] = A[j * n + i] = 0.0;
            if (t % 7 == 0) A[1] = A[n] = 0.0;
            bool a = invert(A.data(), n, X.data()), b = ref_invert(A.data(), n, Y.data());
            if (a != b || std::memcmp(X.data(), Y.data(), sizeof(double) * n * n)) failures++, std::printf("invert n=%d trial %d differs\n", n, t);
            a = solve_first_columns(A.data(), n, m, X.data());
            b = ref_solve_first_columns(A.data(), n, m, Y.data());
            if (a != b || std::memcmp(X.data(), Y.data(), sizeof(double) * n * m)) failures++, std::printf("solve n=%d trial %d differs\n", n, t);
        }
        std::vector<double> Z(n * n, 0.0);  // singular: both must say so
        if (invert(Z.data(), n, X.data()) || solve_first_columns(Z.data(), n, m, X.data())) failures++, std::printf("singular n=%d accepted\n", n);
    }
    // ---- ImuProcess: poses first + FinishCovariance later == everything at once (IMU_Processing.hpp:204-330)
    for (int trial = 0; trial < 20; trial++) {
        ImuProcess now, later;
        ImuSample last;
        last.t = 99.9975;
        for (int i = 0; i < 3; i++) last.acc[i] = (i == 2) ? 9.81 : 0.0, last.gyr[i] = 0.0;
        now.force_ready(Vec3(0.1, -0.2, 9.7), last);
        later.force_ready(Vec3(0.1, -0.2, 9.7), last);
        StatesGroup a, b;
        for (int scan = 0; scan < 6; scan++) {
            std::vector<ImuSample> v;
            const double t0 = 100.0 + scan * 0.1;
            const int n_imu = 18 + (int)(g() % 5);
            for (int i = 0; i < n_imu; i++) {
                ImuSample s;
                s.t = t0 + (i + 0.5) * 0.1 / n_imu - 0.004;
                for (int k = 0; k < 3; k++) s.acc[k] = ((k == 2) ? 9.81 : 0.0) + 0.3 * N(g), s.gyr[k] = 0.2 * N(g);
                v.push_back(s);
            }
            const bool stop = (scan == 4);
            const bool ra = now.Process(v, t0, t0 + 0.1, a, stop);
            const bool rb = later.Process(v, t0, t0 + 0.1, b, stop, /*defer_cov=*/true);
            if (ra != rb) failures++;
            if (now.IMUpose.size() != later.IMUpose.size() ||
                std::memcmp(now.IMUpose.data(), later.IMUpose.data(), now.IMUpose.size() * sizeof(Pose6D)))
                failures++, std::printf("IMUpose differs (trial %d scan %d)\n", trial, scan);
            if (scan % 2 == 0) later.FinishCovariance(b);  // odd scans: left pending, the next Process finishes it first
            double fa[36 + kDim * kDim], fb[36 + kDim * kDim];
            a.to_flat(fa);
            b.to_flat(fb);
            if (std::memcmp(fa, fb, 36 * sizeof(double))) failures++, std::printf("state differs (trial %d scan %d)\n", trial, scan);
            if (scan % 2 == 0 && std::memcmp(fa + 36, fb + 36, kDim * kDim * sizeof(double)))
                failures++, std::printf("covariance differs (trial %d scan %d)\n", trial, scan);
        }
        later.FinishCovariance(b);
        if (std::memcmp(a.cov, b.cov, sizeof(a.cov))) failures++, std::printf("final covariance differs (trial %d)\n", trial);
    }
    std::printf("host_unit: %d failure(s)\n", failures);
    return failures ? 1 : 0;
}
