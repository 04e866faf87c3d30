// daliti_b200/csrc/host/eskf_lio_host.cpp -- see eskf_lio_host.hpp.
// Reference lines are cited as path:line relative to the DaLiTI tree.
#include "eskf_lio_host.hpp"

#include <algorithm>
#include <chrono>

namespace dlt_host {

static double wall() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// ------------------------------------------------------------------ SO(3)
static Mat3 rodrigues(const Vec3 &axis, double ang) {
    // Eye3 + sin(ang) K + (1 - cos(ang)) K K
    Mat3 K = Mat3::hat(axis);
    return Mat3::identity() + K * std::sin(ang) + (K * (1.0 - std::cos(ang))) * K;
}
Mat3 so3_exp_rate(const Vec3 &w, double dt) {
    double n = w.norm();
    if (n > 0.0000001) return rodrigues(w / n, n * dt);
    return Mat3::identity();
}
Mat3 so3_exp(double v1, double v2, double v3) {
    double n = std::sqrt(v1 * v1 + v2 * v2 + v3 * v3);
    if (n > 0.00001) return rodrigues(Vec3(v1 / n, v2 / n, v3 / n), n);
    return Mat3::identity();
}
Vec3 so3_log(const Mat3 &R) {
    double tr = R.trace();
    double theta = (tr > 3.0 - 1e-6) ? 0.0 : std::acos(0.5 * (tr - 1));
    Vec3 K(R.a[7] - R.a[5], R.a[2] - R.a[6], R.a[3] - R.a[1]);
    return (std::abs(theta) < 0.001) ? (K * 0.5) : (K * (0.5 * theta / std::sin(theta)));
}
Mat3 quat_to_rot(double w, double x, double y, double z) {  // Eigen::Quaterniond::toRotationMatrix
    Mat3 r;
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    r.a[0] = 1 - (tyy + tzz);
    r.a[1] = txy - twz;
    r.a[2] = txz + twy;
    r.a[3] = txy + twz;
    r.a[4] = 1 - (txx + tzz);
    r.a[5] = tyz - twx;
    r.a[6] = txz - twy;
    r.a[7] = tyz + twx;
    r.a[8] = 1 - (txx + tyy);
    return r;
}

// ------------------------------------------------------------------ StatesGroup
StatesGroup::StatesGroup() {
    std::memset(cov, 0, sizeof(cov));
    for (int i = 0; i < kDim; i++) cov[i * kDim + i] = 1.0;  // INIT_COV, common_lib.h:28,85
}
StatesGroup StatesGroup::boxplus(const double *d) const {
    StatesGroup a;
    a.rot_end = rot_end * so3_exp(d[0], d[1], d[2]);
    a.pos_end = pos_end + Vec3(d + 3);
    a.R_L_I = R_L_I * so3_exp(d[6], d[7], d[8]);
    a.T_L_I = T_L_I + Vec3(d + 9);
    a.vel_end = vel_end + Vec3(d + 12);
    a.bias_g = bias_g + Vec3(d + 15);
    a.bias_a = bias_a + Vec3(d + 18);
    a.gravity = gravity + Vec3(d + 21);
    std::memcpy(a.cov, cov, sizeof(cov));
    return a;
}
StatesGroup StatesGroup::compose(const StatesGroup &b) const {
    StatesGroup r;
    r.rot_end = rot_end * b.rot_end;
    r.pos_end = pos_end + b.pos_end;
    r.R_L_I = R_L_I * b.R_L_I;
    r.T_L_I = T_L_I + b.T_L_I;
    r.vel_end = vel_end + b.vel_end;
    r.bias_g = bias_g;
    r.bias_a = bias_a;
    r.gravity = gravity;
    std::memcpy(r.cov, cov, sizeof(cov));
    return r;
}
void StatesGroup::boxplus_inplace(const double *d) {
    rot_end = rot_end * so3_exp(d[0], d[1], d[2]);
    pos_end = pos_end + Vec3(d + 3);
    R_L_I = R_L_I * so3_exp(d[6], d[7], d[8]);
    T_L_I = T_L_I + Vec3(d + 9);
    vel_end = vel_end + Vec3(d + 12);
    bias_g = bias_g + Vec3(d + 15);
    bias_a = bias_a + Vec3(d + 18);
    gravity = gravity + Vec3(d + 21);
}
void StatesGroup::boxminus(const StatesGroup &b, double *o) const {
    so3_log(b.rot_end.t() * rot_end).store(o);
    (pos_end - b.pos_end).store(o + 3);
    so3_log(b.R_L_I.t() * R_L_I).store(o + 6);
    (T_L_I - b.T_L_I).store(o + 9);
    (vel_end - b.vel_end).store(o + 12);
    (bias_g - b.bias_g).store(o + 15);
    (bias_a - b.bias_a).store(o + 18);
    (gravity - b.gravity).store(o + 21);
}
StatesGroup StatesGroup::scaled(double s) const {
    StatesGroup a;
    Vec3 so3 = so3_log(rot_end);
    a.rot_end = so3_exp(so3.x * s, so3.y * s, so3.z * s);
    a.pos_end = pos_end * s;
    a.R_L_I = R_L_I;
    a.T_L_I = T_L_I * s;
    a.vel_end = vel_end * s;
    a.bias_g = bias_g * s;
    a.bias_a = bias_a * s;
    a.gravity = gravity * s;
    std::memcpy(a.cov, cov, sizeof(cov));
    return a;
}
void StatesGroup::pose24(double *p) const {
    std::memcpy(p, rot_end.a, 72);
    pos_end.store(p + 9);
    std::memcpy(p + 12, R_L_I.a, 72);
    T_L_I.store(p + 21);
}
void StatesGroup::to_flat(double *f) const {
    pose24(f);
    vel_end.store(f + 24);
    bias_g.store(f + 27);
    bias_a.store(f + 30);
    gravity.store(f + 33);
    std::memcpy(f + 36, cov, sizeof(cov));
}
void StatesGroup::from_flat(const double *f) {
    rot_end = Mat3::from(f);
    pos_end = Vec3(f + 9);
    R_L_I = Mat3::from(f + 12);
    T_L_I = Vec3(f + 21);
    vel_end = Vec3(f + 24);
    bias_g = Vec3(f + 27);
    bias_a = Vec3(f + 30);
    gravity = Vec3(f + 33);
    std::memcpy(cov, f + 36, sizeof(cov));
}

// ------------------------------------------------------------------ dense helpers
// The two 24 x 24 systems of the Kalman update sit on the critical path of the host loop (the GPU idles while they are
// solved), so the sizes the update uses get a fixed-size instance: compile-time bounds, a stack work area, and -- on
// x86-64 -- an AVX2 clone picked at load time.  Same operations in the same order as the generic routine (no FMA: the
// clone enables AVX2 only), so the results are bit-identical; 8 -> 2.6 us per solve on the build machine.
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__) && !defined(DLT_NO_TARGET_CLONES)
#define DLT_HOST_CLONES __attribute__((target_clones("avx2", "default")))
#else
#define DLT_HOST_CLONES
#endif

// Gauss-Jordan on [A | I] with partial pivoting
template <int N>
__attribute__((always_inline)) static inline bool invert_fixed(const double *A, double *out) {
    constexpr int W = 2 * N;
    alignas(32) double w[N][W];
    for (int i = 0; i < N; i++) {
        for (int j = 0; j < N; j++) {
            w[i][j] = A[i * N + j];
            w[i][N + j] = (i == j) ? 1.0 : 0.0;
        }
    }
    for (int c = 0; c < N; c++) {
        int p = c;
        for (int r = c + 1; r < N; r++)
            if (std::fabs(w[r][c]) > std::fabs(w[p][c])) p = r;
        if (w[p][c] == 0.0) return false;
        if (p != c)
            for (int j = 0; j < W; j++) std::swap(w[p][j], w[c][j]);
        const double inv = 1.0 / w[c][c];
        for (int j = 0; j < W; j++) w[c][j] *= inv;
        for (int r = 0; r < N; r++) {
            if (r == c) continue;
            const double f = w[r][c];
            if (f == 0.0) continue;
            for (int j = 0; j < W; j++) w[r][j] -= f * w[c][j];
        }
    }
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) out[i * N + j] = w[i][N + j];
    return true;
}
DLT_HOST_CLONES static bool invert24(const double *A, double *out) { return invert_fixed<24>(A, out); }

bool invert(const double *A, int n, double *out) {
    if (n == 24) return invert24(A, out);
    std::vector<double> w((size_t)n * 2 * n);
    const int W = 2 * n;
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < n; j++) {
            w[(size_t)i * W + j] = A[(size_t)i * n + j];
            w[(size_t)i * W + n + j] = (i == j) ? 1.0 : 0.0;
        }
    }
    for (int c = 0; c < n; c++) {
        int p = c;
        for (int r = c + 1; r < n; r++)
            if (std::fabs(w[(size_t)r * W + c]) > std::fabs(w[(size_t)p * W + c])) p = r;
        if (w[(size_t)p * W + c] == 0.0) return false;
        if (p != c)
            for (int j = 0; j < W; j++) std::swap(w[(size_t)p * W + j], w[(size_t)c * W + j]);
        const double inv = 1.0 / w[(size_t)c * W + c];
        for (int j = 0; j < W; j++) w[(size_t)c * W + j] *= inv;
        for (int r = 0; r < n; r++) {
            if (r == c) continue;
            const double f = w[(size_t)r * W + c];
            if (f == 0.0) continue;
            for (int j = 0; j < W; j++) w[(size_t)r * W + j] -= f * w[(size_t)c * W + j];
        }
    }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) out[(size_t)i * n + j] = w[(size_t)i * W + n + j];
    return true;
}

// X (n x m, row-major) = first m columns of A^-1, by LU with partial pivoting on [A | I[:, :m]]
template <int N, int M>
__attribute__((always_inline)) static inline bool solve_first_columns_fixed(const double *A, double *X) {
    constexpr int W = N + M;
    alignas(32) double w[N][W];
    for (int i = 0; i < N; i++) {
        for (int j = 0; j < N; j++) w[i][j] = A[i * N + j];
        for (int j = 0; j < M; j++) w[i][N + j] = (i == j) ? 1.0 : 0.0;
    }
    for (int c = 0; c < N; c++) {
        int p = c;
        for (int r = c + 1; r < N; r++)
            if (std::fabs(w[r][c]) > std::fabs(w[p][c])) p = r;
        if (w[p][c] == 0.0) return false;
        if (p != c)
            for (int j = c; j < W; j++) std::swap(w[p][j], w[c][j]);
        const double piv = w[c][c];
        for (int r = c + 1; r < N; r++) {
            const double f = w[r][c] / piv;
            if (f == 0.0) continue;
            for (int j = c + 1; j < W; j++) w[r][j] -= f * w[c][j];
        }
    }
    for (int c = N - 1; c >= 0; c--) {  // all M right-hand sides side by side: per column the same k order as below
        const double piv = w[c][c];
        double s[M];
        for (int j = 0; j < M; j++) s[j] = w[c][N + j];
        for (int k = c + 1; k < N; k++) {
            const double a = w[c][k];
            for (int j = 0; j < M; j++) s[j] -= a * X[k * M + j];
        }
        for (int j = 0; j < M; j++) X[c * M + j] = s[j] / piv;
    }
    return true;
}
DLT_HOST_CLONES static bool solve_first_columns_24_12(const double *A, double *X) { return solve_first_columns_fixed<24, 12>(A, X); }

bool solve_first_columns(const double *A, int n, int m, double *X) {
    if (n == 24 && m == 12) return solve_first_columns_24_12(A, X);
    const int W = n + m;
    std::vector<double> w((size_t)n * W);
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < n; j++) w[(size_t)i * W + j] = A[(size_t)i * n + j];
        for (int j = 0; j < m; j++) w[(size_t)i * W + n + j] = (i == j) ? 1.0 : 0.0;
    }
    for (int c = 0; c < n; c++) {
        int p = c;
        for (int r = c + 1; r < n; r++)
            if (std::fabs(w[(size_t)r * W + c]) > std::fabs(w[(size_t)p * W + c])) p = r;
        if (w[(size_t)p * W + c] == 0.0) return false;
        if (p != c)
            for (int j = c; j < W; j++) std::swap(w[(size_t)p * W + j], w[(size_t)c * W + j]);
        const double piv = w[(size_t)c * W + c];
        for (int r = c + 1; r < n; r++) {
            const double f = w[(size_t)r * W + c] / piv;
            if (f == 0.0) continue;
            for (int j = c + 1; j < W; j++) w[(size_t)r * W + j] -= f * w[(size_t)c * W + j];
        }
    }
    for (int c = n - 1; c >= 0; c--) {
        const double piv = w[(size_t)c * W + c];
        for (int j = 0; j < m; j++) {
            double s = w[(size_t)c * W + n + j];
            for (int k = c + 1; k < n; k++) s -= w[(size_t)c * W + k] * X[(size_t)k * m + j];
            X[(size_t)c * m + j] = s / piv;
        }
    }
    return true;
}

static void set_block3(double *M, int r0, int c0, const Mat3 &B) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) M[(r0 + i) * kDim + c0 + j] = B.a[3 * i + j];
}

// ------------------------------------------------------------------ ImuProcess
static const double G_m_s2 = 9.8099;  // common_lib.h:23
static const int MAX_INI_COUNT = 100;  // IMU_Processing.hpp:29

ImuProcess::ImuProcess() { std::memset(&last_imu_, 0, sizeof(last_imu_)); }

void ImuProcess::Reset() {
    angvel_last = Vec3();
    cov_acc = Vec3(0.1, 0.1, 0.1);
    cov_gyr = Vec3(0.1, 0.1, 0.1);
    mean_acc = Vec3(0, 0, -1.0);
    mean_gyr = Vec3();
    imu_need_init_ = true;
    b_first_frame_ = true;
    init_iter_num = 1;
    std::memset(&last_imu_, 0, sizeof(last_imu_));
    IMUpose.clear();
}
void ImuProcess::set_extrinsic(const Vec3 &t, const Mat3 &r) {
    Lidar_T_wrt_IMU = t;
    Lidar_R_wrt_IMU = r;
}
void ImuProcess::force_ready(const Vec3 &m, const ImuSample &last) {
    imu_need_init_ = false;
    b_first_frame_ = false;
    init_iter_num = MAX_INI_COUNT + 1;
    mean_acc = m;
    last_imu_ = last;
}

void ImuProcess::IMU_Initial(const std::vector<ImuSample> &imu, StatesGroup &st, int &N) {
    if (b_first_frame_) {
        Reset();
        N = 1;
        b_first_frame_ = false;
        mean_acc = Vec3(imu.front().acc);
        mean_gyr = Vec3(imu.front().gyr);
    }
    for (const ImuSample &s : imu) {
        Vec3 ca(s.acc), cg(s.gyr);
        mean_acc = mean_acc + (ca - mean_acc) / N;
        mean_gyr = mean_gyr + (cg - mean_gyr) / N;
        Vec3 da = ca - mean_acc, dg = cg - mean_gyr;
        cov_acc = Vec3(cov_acc.x * (N - 1.0) / N + da.x * da.x * (N - 1.0) / (N * N), cov_acc.y * (N - 1.0) / N + da.y * da.y * (N - 1.0) / (N * N),
                       cov_acc.z * (N - 1.0) / N + da.z * da.z * (N - 1.0) / (N * N));
        cov_gyr = Vec3(cov_gyr.x * (N - 1.0) / N + dg.x * dg.x * (N - 1.0) / (N * N), cov_gyr.y * (N - 1.0) / N + dg.y * dg.y * (N - 1.0) / (N * N),
                       cov_gyr.z * (N - 1.0) / N + dg.z * dg.z * (N - 1.0) / (N * N));
        N++;
    }
    st.gravity = (mean_acc * -1.0) / mean_acc.norm() * G_m_s2;
    st.bias_g = mean_gyr;
    st.R_L_I = Lidar_R_wrt_IMU;
    st.T_L_I = Lidar_T_wrt_IMU;
    last_imu_ = imu.back();
}

// One IMU step of cov = F cov F^T + Q, IMU_Processing.hpp:262-288.  F = I + six 3x3 blocks, Q = five 3x3 diagonal blocks
// (:270-286): the caller's F / Q buffers keep the fixed structure, only the block values are rewritten here.
void ImuProcess::PropagateCovStep(const CovStep &step, StatesGroup &st, double *__restrict__ F, double *__restrict__ Q, double *__restrict__ FP) const {
    const Vec3 &angvel_avr = step.angvel_avr, &acc_avr = step.acc_avr;
    const Mat3 &R_imu = step.R_imu;
    const double dt = step.dt;
    set_block3(F, 0, 0, so3_exp_rate(angvel_avr, -dt));
    set_block3(F, 0, 15, Mat3::identity() * (-dt));
    set_block3(F, 3, 12, Mat3::identity() * dt);
    set_block3(F, 12, 0, (R_imu * -1.0) * Mat3::hat(acc_avr) * dt);
    set_block3(F, 12, 18, R_imu * (-dt));
    set_block3(F, 12, 21, Mat3::identity() * dt);
    Mat3 Ca, Cg;
    Ca.a[0] = cov_acc.x;
    Ca.a[4] = cov_acc.y;
    Ca.a[8] = cov_acc.z;
    Cg.a[0] = cov_gyr.x;
    Cg.a[4] = cov_gyr.y;
    Cg.a[8] = cov_gyr.z;
    Q[0 * kDim + 0] = cov_gyr.x * dt * dt * 10000;
    Q[1 * kDim + 1] = cov_gyr.y * dt * dt * 10000;
    Q[2 * kDim + 2] = cov_gyr.z * dt * dt * 10000;
    set_block3(Q, 3, 3, R_imu * Cg * R_imu.t() * dt * dt * 10000);
    set_block3(Q, 12, 12, R_imu * Ca * R_imu.t() * dt * dt * 10000);
    for (int a = 0; a < 3; a++) {
        Q[(15 + a) * kDim + 15 + a] = 0.0001 * dt * dt;
        Q[(18 + a) * kDim + 18 + a] = 0.0001 * dt * dt;
    }
    // cov = F cov F^T + Q with the block structure written out (terms in ascending column order of F, i.e. the
    // order a dense product would add the non-zeros in).  Rows / columns 0-5 and 12-14 are the only non-identity ones.
    {
        const double *E = &F[0];            // F[0:3, 0:3]   = Exp(w, -dt)
        const double *A = &F[12 * kDim];    // F[12:15, 0:3] = -R a^ dt   (row stride kDim)
        const double *B = &F[12 * kDim + 18];  // F[12:15, 18:21] = -R dt
        const double *P = st.cov;
        std::memcpy(FP, P, sizeof(double) * kDim * kDim);
        for (int r = 0; r < 3; r++) {
            double *o0 = &FP[r * kDim], *o3 = &FP[(3 + r) * kDim], *o12 = &FP[(12 + r) * kDim];
            const double e0 = E[r * kDim], e1 = E[r * kDim + 1], e2 = E[r * kDim + 2];
            const double a0 = A[r * kDim], a1 = A[r * kDim + 1], a2 = A[r * kDim + 2];
            const double b0 = B[r * kDim], b1 = B[r * kDim + 1], b2 = B[r * kDim + 2];
            for (int j = 0; j < kDim; j++) {
                o0[j] = ((e0 * P[j] + e1 * P[kDim + j]) + e2 * P[2 * kDim + j]) + (-dt) * P[(15 + r) * kDim + j];
                o3[j] = P[(3 + r) * kDim + j] + dt * P[(12 + r) * kDim + j];
                o12[j] = ((((((a0 * P[j] + a1 * P[kDim + j]) + a2 * P[2 * kDim + j]) + P[(12 + r) * kDim + j]) + b0 * P[18 * kDim + j]) +
                           b1 * P[19 * kDim + j]) + b2 * P[20 * kDim + j]) + dt * P[(21 + r) * kDim + j];
            }
        }
        for (int i = 0; i < kDim; i++) {
            const double *X = &FP[i * kDim];
            double *o = &st.cov[i * kDim];
            double c0[3], c3[3], c12[3];
            for (int c = 0; c < 3; c++) {
                c0[c] = ((E[c * kDim] * X[0] + E[c * kDim + 1] * X[1]) + E[c * kDim + 2] * X[2]) + (-dt) * X[15 + c];
                c3[c] = X[3 + c] + dt * X[12 + c];
                c12[c] = ((((((A[c * kDim] * X[0] + A[c * kDim + 1] * X[1]) + A[c * kDim + 2] * X[2]) + X[12 + c]) + B[c * kDim] * X[18]) +
                           B[c * kDim + 1] * X[19]) + B[c * kDim + 2] * X[20]) + dt * X[21 + c];
            }
            for (int j = 0; j < kDim; j++) o[j] = X[j] + Q[i * kDim + j];
            for (int c = 0; c < 3; c++) {
                o[c] = c0[c] + Q[i * kDim + c];
                o[3 + c] = c3[c] + Q[i * kDim + 3 + c];
                o[12 + c] = c12[c] + Q[i * kDim + 12 + c];
            }
        }
    }
}

void ImuProcess::FinishCovariance(StatesGroup &st) {
    if (pending_cov_.empty()) return;
    double F[kDim * kDim], Q[kDim * kDim], FP[kDim * kDim];
    std::memset(F, 0, sizeof(F));
    std::memset(Q, 0, sizeof(Q));
    for (int i = 0; i < kDim; i++) F[i * kDim + i] = 1.0;
    for (const CovStep &cs : pending_cov_) PropagateCovStep(cs, st, F, Q, FP);
    pending_cov_.clear();
}

void ImuProcess::Propagate(const std::vector<ImuSample> &imu, double pcl_beg_time, double pcl_end_time, StatesGroup &st, bool EKF_stop_flg,
                           bool defer_cov) {
    std::vector<ImuSample> v;
    v.reserve(imu.size() + 1);
    v.push_back(last_imu_);  // IMU_Processing.hpp:207-208
    v.insert(v.end(), imu.begin(), imu.end());
    const double imu_end_time = v.back().t;

    IMUpose.clear();
    Pose6D p0;
    p0.offset_time = 0.0;
    acc_s_last.store(p0.acc);
    angvel_last.store(p0.gyr);
    st.vel_end.store(p0.vel);
    st.pos_end.store(p0.pos);
    std::memcpy(p0.rot, st.rot_end.a, sizeof(p0.rot));
    IMUpose.push_back(p0);  // :224

    Vec3 acc_imu, angvel_avr, acc_avr, vel_imu = st.vel_end, pos_imu = st.pos_end;
    Mat3 R_imu = st.rot_end;
    double dt = 0;
    // F = I + six 3x3 blocks, Q = five 3x3 diagonal blocks (:270-286): the structure is fixed, only the block
    // values change per IMU step, so the buffers are set up once and the products only visit the non-zeros
    double F[kDim * kDim], Q[kDim * kDim], FP[kDim * kDim];
    std::memset(F, 0, sizeof(F));
    std::memset(Q, 0, sizeof(Q));
    for (int i = 0; i < kDim; i++) F[i * kDim + i] = 1.0;
    for (size_t k = 0; k + 1 < v.size(); k++) {
        const ImuSample &head = v[k], &tail = v[k + 1];
        if (tail.t < last_observation_end_time_) continue;  // :237
        angvel_avr = (Vec3(head.gyr) + Vec3(tail.gyr)) * 0.5 - st.bias_g;
        acc_avr = (Vec3(head.acc) + Vec3(tail.acc)) * 0.5 * G_m_s2 / mean_acc.norm() - st.bias_a;  // :248
        dt = (head.t < last_observation_end_time_) ? tail.t - last_observation_end_time_ : tail.t - head.t;

        const Mat3 Exp_f = so3_exp_rate(angvel_avr, dt);
        {  // covariance propagation, :262-288 (now, or recorded for FinishCovariance)
            CovStep cs;
            cs.angvel_avr = angvel_avr;
            cs.acc_avr = acc_avr;
            cs.R_imu = R_imu;
            cs.dt = dt;
            if (defer_cov)
                pending_cov_.push_back(cs);
            else
                PropagateCovStep(cs, st, F, Q, FP);
        }

        R_imu = R_imu * Exp_f;                                      // :291
        acc_imu = R_imu * acc_avr + st.gravity;                     // :294
        pos_imu = pos_imu + vel_imu * dt + acc_imu * 0.5 * dt * dt;  // :297
        vel_imu = vel_imu + acc_imu * dt;                           // :300
        angvel_last = angvel_avr;
        acc_s_last = acc_imu;
        Pose6D p;
        p.offset_time = tail.t - pcl_beg_time;
        acc_imu.store(p.acc);
        angvel_avr.store(p.gyr);
        vel_imu.store(p.vel);
        pos_imu.store(p.pos);
        std::memcpy(p.rot, R_imu.a, sizeof(p.rot));
        IMUpose.push_back(p);  // :307
    }
    dt = pcl_end_time - imu_end_time;  // :313
    if (!EKF_stop_flg) {
        st.vel_end = vel_imu + acc_imu * dt;
        st.rot_end = R_imu * so3_exp_rate(angvel_avr, dt);
        st.pos_end = pos_imu + vel_imu * dt + acc_imu * 0.5 * dt * dt;
    }
    last_observation_end_time_ = pcl_end_time;  // :325
}

bool ImuProcess::Process(const std::vector<ImuSample> &imu, double lidar_beg_time, double observation_end_time, StatesGroup &st,
                         bool EKF_stop_flg, bool defer_cov) {
    if (imu.empty()) return false;  // :378-382
    if (imu_need_init_) {
        IMU_Initial(imu, st, init_iter_num);
        imu_need_init_ = true;
        last_imu_ = imu.back();
        if (init_iter_num > MAX_INI_COUNT) {
            imu_need_init_ = false;
            cov_acc = Vec3(0.1, 0.1, 0.1);
            cov_gyr = Vec3(0.1, 0.1, 0.1);
        }
        return false;
    }
    st.gravity = Vec3(0, 0, -9.801);  // :408-412: overwritten by constants every scan (reference quirk)
    st.bias_a = Vec3(1.44e-04, 1.44e-04, 1.44e-04);
    st.bias_g = Vec3(5.3e-05, 5.3e-05, 5.3e-05);
    st.R_L_I = Lidar_R_wrt_IMU;
    st.T_L_I = Lidar_T_wrt_IMU;
    FinishCovariance(st);  // (a caller that deferred and never finished: keep the sequence of updates intact)
    Propagate(imu, lidar_beg_time, observation_end_time, st, EKF_stop_flg, defer_cov);
    last_imu_ = imu.back();  // :422
    return true;
}

// ------------------------------------------------------------------ LaserMapping
static StatesGroup odom_to_state(const double *pos, const double *q, const double *vel, const double *cs) {
    StatesGroup r;  // odomToStateGruop, laserMapping.cpp:218-239
    r.pos_end = Vec3(pos);
    r.rot_end = quat_to_rot(q[0], q[1], q[2], q[3]);
    r.vel_end = Vec3(vel);
    r.bias_g = Vec3(cs[4], cs[5], cs[6]);
    r.bias_a = Vec3(cs[1], cs[2], cs[3]);
    r.gravity = Vec3(0, 0, cs[7]);
    return r;
}

LaserMapping::LaserMapping(const dlt_lio_config &c) : cfg(c) {
    imu_.set_extrinsic(Vec3(cfg.extrinT), Mat3::from(cfg.extrinR));  // laserMapping.cpp:666-667, 706
    create_rc_ = dlt_create(&cfg.dev, &dev_);
    if (create_rc_ != 0) dev_ = nullptr;
}
LaserMapping::~LaserMapping() {
    if (dev_) dlt_destroy(dev_);
    delete iekf_blk_;
}
void LaserMapping::on_lidar_msg() {
    lidar_frame_counter_num++;
    if (lidar_frame_counter_num > 100) dynamic_effect_featurepoints_threshold = cfg.featptsThreshold;
    lidar_cnt++;
}
void LaserMapping::on_edge_count(int recv_n) {
    double alpha_t = recv_n / 307200.0;  // Mta, laserMapping.cpp:107
    zeta_t = 2.0 / (1.0 + std::exp(-alpha_t)) - 1;
}

int LaserMapping::fov_segment(const Vec3 &pos, int *deleted) {
    *deleted = 0;
    const float MOV_THRESHOLD = 1.5f, DET_RANGE = cfg.det_range;
    const double p[3] = {pos.x, pos.y, pos.z};
    if (!Localmap_Initialized) {
        for (int i = 0; i < 3; i++) {
            LocalMap_min[i] = (float)(p[i] - cfg.cube_len / 2.0);
            LocalMap_max[i] = (float)(p[i] + cfg.cube_len / 2.0);
        }
        Localmap_Initialized = true;
        return 0;
    }
    float edge[3][2];
    bool need_move = false;
    for (int i = 0; i < 3; i++) {
        edge[i][0] = (float)std::fabs(p[i] - (double)LocalMap_min[i]);
        edge[i][1] = (float)std::fabs(p[i] - (double)LocalMap_max[i]);
        if (edge[i][0] <= MOV_THRESHOLD * DET_RANGE || edge[i][1] <= MOV_THRESHOLD * DET_RANGE) need_move = true;
    }
    if (!need_move) return 0;
    float nmin[3], nmax[3];
    std::memcpy(nmin, LocalMap_min, sizeof(nmin));
    std::memcpy(nmax, LocalMap_max, sizeof(nmax));
    const float mov_dist = (float)std::max((cfg.cube_len - 2.0 * MOV_THRESHOLD * DET_RANGE) * 0.5 * 0.9, double(DET_RANGE * (MOV_THRESHOLD - 1)));
    std::vector<float> boxes;
    for (int i = 0; i < 3; i++) {
        float bmin[3], bmax[3];
        std::memcpy(bmin, LocalMap_min, sizeof(bmin));
        std::memcpy(bmax, LocalMap_max, sizeof(bmax));
        if (edge[i][0] <= MOV_THRESHOLD * DET_RANGE) {
            nmax[i] -= mov_dist;
            nmin[i] -= mov_dist;
            bmin[i] = LocalMap_max[i] - mov_dist;
        } else if (edge[i][1] <= MOV_THRESHOLD * DET_RANGE) {
            nmax[i] += mov_dist;
            nmin[i] += mov_dist;
            bmax[i] = LocalMap_min[i] + mov_dist;
        } else {
            continue;
        }
        boxes.insert(boxes.end(), bmin, bmin + 3);
        boxes.insert(boxes.end(), bmax, bmax + 3);
    }
    std::memcpy(LocalMap_min, nmin, sizeof(nmin));
    std::memcpy(LocalMap_max, nmax, sizeof(nmax));
    if (!boxes.empty() && map_built) return dlt_map_delete_boxes(dev_, boxes.data(), (int)(boxes.size() / 6), deleted);
    return 0;
}

// zeta blend, laserMapping.cpp:1105-1131
void LaserMapping::zeta_blend(int effct_feat_num, const StatesGroup &state_propagat, const dlt_lio_thermal *th) {
    if ((lidar_cnt < 100) || (!th->tis_online) || (th->tis_online && lidar_cnt % 2 == 1)) {
        const double alpha_l = effct_feat_num / (cfg.beta * 65536.0);  // Nla, :106
        zeta_l = 2.0 / (1.0 + std::exp(-alpha_l)) - 1;
        double zeta_l_norm = zeta_l / (zeta_l + zeta_t);
        if (lidar_cnt < 100) zeta_l_norm = 1;
        double v1[kDim], v2[kDim];
        state_propagat.boxminus(last_state, v1);
        state.boxminus(last_state, v2);
        for (int i = 0; i < kDim; i++) {
            v1[i] *= (1 - zeta_l_norm);
            v2[i] *= zeta_l_norm;
        }
        state = last_state.boxplus(v1).boxplus(v2);  // :1119 -- the result carries last_state.cov (reference quirk)
    } else {
        const double zeta_t_norm = zeta_t / (zeta_l + zeta_t);
        double v1[kDim];
        state.boxminus(last_state, v1);
        for (int i = 0; i < kDim; i++) v1[i] *= (1 - zeta_t_norm);
        StatesGroup o = odom_to_state(th->l2l_pos, th->l2l_quat, th->l2l_vel, th->l2l_cov_slots);
        state = last_state.boxplus(v1).compose(o.scaled(zeta_t_norm));  // :1127
    }
    last_state = state;  // :1131
}

int LaserMapping::reduce_trampoline(void *self, double *result_dev, int n) {
    LaserMapping *lm = static_cast<LaserMapping *>(self);
    return lm->reduce_fn(lm->reduce_ctx, result_dev, n);
}

#define LM_CK(expr)                               \
    do {                                          \
        int rc__ = (expr);                        \
        if (rc__ != 0) {                          \
            err = dlt_last_error(dev_);           \
            return rc__;                          \
        }                                         \
    } while (0)

int LaserMapping::collect_insert(int *n_ds, int *n_raw) {
    int a = last_added_ds, b = last_added_raw;
    if (insert_pending) {
        insert_pending = false;
        LM_CK(dlt_map_incremental_collect(dev_, &a, &b));
        last_added_ds = a;
        last_added_raw = b;
    }
    if (n_ds) *n_ds = a;
    if (n_raw) *n_raw = b;
    return 0;
}

int LaserMapping::process_scan(const void *pts48, int n, double lidar_beg_time, const ImuSample *imu, int n_imu, const dlt_lio_thermal *th,
                               dlt_lio_scan_out *out, bool pts_on_device, double observation_end_time_in) {
    const double LASER_POINT_COV = 0.0015;  // laserMapping.cpp:76
    const int QUEUE_SIZE = 10;              // :192
    const int NUM_MAX_ITERATIONS = cfg.max_iteration;
    static const dlt_lio_thermal no_thermal = {0, 0, {0, 0, 0}, {1, 0, 0, 0}, {0, 0, 0}, {0, 0, 0, 0, 0, 0, 0, 0},
                                               {0, 0, 0}, {1, 0, 0, 0}, {0, 0, 0}, {0, 0, 0, 0, 0, 0, 0, 0}};
    if (!th) th = &no_thermal;
    std::memset(out, 0, sizeof(*out));
    if (insert_pending) {  // adopt the previous scan's map_incremental (errors of that insert surface here)
        int a = 0, b = 0;
        if (int rc = collect_insert(&a, &b)) return rc;
    }
    iters.clear();
    const double t_begin = wall();
    if (flg_first_scan) {  // :736-740
        first_lidar_time = lidar_beg_time;
        flg_first_scan = false;
    }
    // observation_end_time = lidar_beg_time + points.back().normal_z      :546
    double observation_end_time = lidar_beg_time;
    if (pts_on_device)
        observation_end_time = observation_end_time_in;
    else if (n > 0)
        observation_end_time += (double)reinterpret_cast<const float *>(pts48)[(size_t)(n - 1) * 12 + 6];

    // ---- p_imu->Process(Measures, state, feats_undistort, EKF_stop_flg)   :750
    double t0 = wall();
    std::vector<ImuSample> imu_v(imu, imu + n_imu);
    // the deskew kernel only waits for the IMU poses: the covariance propagation of the same IMU steps (~15 us) runs below,
    // once the scan's first kernels are in flight (finish_cov), instead of in front of them
    bool undistorted = imu_.Process(imu_v, lidar_beg_time, observation_end_time, state, EKF_stop_flg, /*defer_cov=*/true);
    struct FinishCov {  // ... and on every way out of this function
        ImuProcess &imu;
        StatesGroup &st;
        ~FinishCov() { imu.FinishCovariance(st); }
    } finish_cov_guard{imu_, state};
    int n_raw = 0;
    if (undistorted && n > 0) {
        double pose[24];
        state.pose24(pose);
        if (pts_on_device)
            LM_CK(dlt_scan_deskew_dev(dev_, pts48, n, reinterpret_cast<const double *>(imu_.IMUpose.data()), (int)imu_.IMUpose.size(), pose));
        else
            LM_CK(dlt_scan_deskew(dev_, pts48, n, reinterpret_cast<const double *>(imu_.IMUpose.data()), (int)imu_.IMUpose.size(), pose));
        n_raw = n;
    }
    out->t_deskew = wall() - t0;
    StatesGroup state_propagat = state;                                 // :752
    Vec3 pos_lid = state_propagat.pos_end + state_propagat.rot_end * state_propagat.T_L_I;  // :753
    {
        double f[36 + kDim * kDim];
        state_propagat.to_flat(f);
        std::memcpy(out->state_prop, f, sizeof(out->state_prop));
    }
    out->n_raw = n_raw;
    if (n_raw == 0) {  // :755-759  "No point, not ready for odometry, skip this scan"
        out->t_total = wall() - t_begin;
        return 0;
    }
    out->had_points = 1;
    flg_EKF_inited = true;  // :762 with INIT_TIME == 0 (:75)

    t0 = wall();
    LM_CK(fov_segment(pos_lid, &out->deleted));  // :772
    out->t_delete = wall() - t0;

    if (!map_built) {  // a map built directly through the device handle (dlt_map_build) also counts as a root
        int c = 0;
        LM_CK(dlt_map_valid_count(dev_, &c));
        if (c > 0) map_built = true;
    }
    // the device-resident loop needs no host copy of feats_down_size before it has run
    // -1 (default) = by measurement: the host loop for a single-GPU map (same latency as the device loop within noise, fewer
    // launches per scan: +46 % aggregate scans/s when many sequences share a GPU), the device loop for a sharded map (the
    // all-reduce then stays in-stream and the host synchronises twice per scan instead of once per iteration)
    // (measured on 2 B200s, C4: host loop + peer mailboxes 0.313 ms p50, device loop 0.371 ms; with a reduce CALLBACK the device loop
    //  keeps the library collective in-stream, 0.375 ms against 0.377 ms + a long tail for the host loop)
    const int loop_mode = cfg.device_loop >= 0 ? cfg.device_loop : ((reduce_fn && !peers) ? 2 : 0);
    const bool device_loop = loop_mode != 0 && map_built && NUM_MAX_ITERATIONS >= 1 && NUM_MAX_ITERATIONS <= DLT_IEKF_MAX_ITER;
    t0 = wall();
    int feats_down_size = 0;
    // (with a map in place nobody needs feats_down_size on the host before the first evaluation of the measurement model
    // has come back, which brings it along: no synchronisation is spent on the VoxelGrid)
    if (device_loop || (map_built && (!reduce_fn || peers)))
        LM_CK(dlt_scan_downsample_async(dev_));
    else
        LM_CK(dlt_scan_downsample(dev_, &feats_down_size));  // :775-778
    out->t_voxel = wall() - t0;
    out->n_down = feats_down_size;
    imu_.FinishCovariance(state);  // deskew + VoxelGrid are running: now the covariance propagation (state.cov is read from here on)
    std::memcpy(state_propagat.cov, state.cov, sizeof(state.cov));

    if (!map_built) {  // ikdtree.Root_Node == nullptr, :780-793
        if (feats_down_size > 5) {
            double pose[24];
            state.pose24(pose);
            LM_CK(dlt_map_build_from_scan(dev_, pose));
            map_built = true;
            out->built_map = 1;
        }
        out->t_total = wall() - t_begin;
        return 0;
    }
    int featsFromMapNum = 0;
    LM_CK(dlt_map_valid_count(dev_, &featsFromMapNum));  // :794
    out->map_points_before = featsFromMapNum;

    int effct_feat_num = 0;
    if (featsFromMapNum < 5) {  // no update: still report feats_down_size
        int nd = 0;
        LM_CK(dlt_scan_get_down(dev_, nullptr, 0, &nd));
        out->n_down = nd;
    }
    if (featsFromMapNum >= 5 && device_loop) {  // :804, the loop :820-1102 resident on the device
        out->did_update = 1;
        t0 = wall();
        if (!iekf_blk_) iekf_blk_ = new dlt_iekf_block();
        dlt_iekf_block &B = *iekf_blk_;
        {
            double f[36 + kDim * kDim];
            state_propagat.to_flat(f);
            std::memcpy(B.state_propagat, f, sizeof(B.state_propagat));
            StatesGroup d = odom_to_state(th->delta_pos, th->delta_quat, th->delta_vel, th->cov_slots);
            d.to_flat(f);
            std::memcpy(B.thermal_delta, f, sizeof(B.thermal_delta));
            B.laser_point_cov = LASER_POINT_COV;
            state.to_flat(B.state);
            last_nodegared_state.to_flat(B.last_nodegared);
        }
        B.threshold = dynamic_effect_featurepoints_threshold;
        B.max_iteration = NUM_MAX_ITERATIONS;
        // zeta blend + map_incremental behind the loop on the device (one synchronisation per scan)
        B.finish = loop_mode == 2 ? 0 : 1;  // 2: loop on the device, blend + insert driven by the host
        B.blend_mode = ((lidar_cnt < 100) || (!th->tis_online) || (th->tis_online && lidar_cnt % 2 == 1)) ? 1 : 2;  // :1107
        {
            double f[36 + kDim * kDim];
            last_state.to_flat(f);
            std::memcpy(B.last_state, f, sizeof(B.last_state));
            StatesGroup o = odom_to_state(th->l2l_pos, th->l2l_quat, th->l2l_vel, th->l2l_cov_slots);
            o.to_flat(f);
            std::memcpy(B.l2l_state, f, sizeof(B.l2l_state));
        }
        B.zeta_t = zeta_t;
        B.zeta_l = zeta_l;
        B.beta = cfg.beta;
        B.lidar_cnt_lt_100 = lidar_cnt < 100 ? 1 : 0;
        B.far_enqueued = 0;
        B.queue_len = (int)effct_q.size();
        for (int i = 0; i < 10; i++) B.effct_queue[i] = i < B.queue_len ? effct_q[i] : 0;
        B.flg_EKF_inited = flg_EKF_inited ? 1 : 0;
        const bool use_cb = reduce_fn && !peers;  // (with peer mailboxes the sum over the ranks happens inside k_residual)
        LM_CK(dlt_iekf_update(dev_, &B, use_cb ? &LaserMapping::reduce_trampoline : nullptr, this, use_cb ? reduce_buf_dev : nullptr));
        // mirror the device's bookkeeping back into the host members
        state.from_flat(B.state);
        last_nodegared_state.from_flat(B.last_nodegared);
        effct_q.assign(B.effct_queue, B.effct_queue + B.queue_len);
        flg_EKF_inited = B.flg_EKF_inited != 0;
        if (B.n_iters > 0) EKF_stop_flg = B.ekf_stop != 0;
        out->n_down = B.n_down;
        for (int k = 0; k < B.n_iters; k++) {
            const dlt_iekf_iter &r = B.iters[k];
            dlt_lio_iter rec;
            std::memset(&rec, 0, sizeof(rec));
            rec.iter = r.iter;
            rec.effct_feat_num = r.effct_feat_num;
            rec.converged = r.converged;
            rec.ekf_stop = r.ekf_stop;
            rec.did_match = r.did_match;
            rec.n_down = B.n_down;
            rec.total_residual = r.total_residual;
            rec.res_mean_last = r.total_residual / r.effct_feat_num;  // :932
            std::memcpy(rec.HtH, r.HtH, sizeof(rec.HtH));
            std::memcpy(rec.Htr, r.Htr, sizeof(rec.Htr));
            std::memcpy(rec.pose_in, r.pose_in, sizeof(rec.pose_in));
            std::memcpy(rec.state_out, r.state_out, sizeof(rec.state_out));
            std::memcpy(rec.solution, r.solution, sizeof(rec.solution));
            iters.push_back(rec);
            effct_feat_num = r.effct_feat_num;
        }
        out->t_iterate = wall() - t0;
        const bool want_eig = B.n_iters > 0;
        t0 = wall();
        if (B.finish && B.n_iters > 0 && B.done) {  // the blend ran on the device: adopt it (covariance of last_state, :1119)
            if (B.blend_mode == 1) zeta_l = B.zeta_l;
            double f[36 + kDim * kDim];
            last_state.to_flat(f);
            std::memcpy(f, B.blend_state, sizeof(B.blend_state));
            state.from_flat(f);
            last_state = state;  // :1131
            if (B.insert_status == 1) {
                out->n_added_ds = B.n_added_ds;
                out->n_added_raw = B.n_added_raw;
                out->added = out->n_added_ds + out->n_added_raw;
            } else if (B.insert_status == 2) {  // more unresolved queries than one fallback chunk: the chunked host-driven path
                double pose[24];
                state.pose24(pose);
                LM_CK(dlt_map_incremental(dev_, pose, flg_EKF_inited ? 1 : 0, &out->n_added_ds, &out->n_added_raw));
                out->added = out->n_added_ds + out->n_added_raw;
            }
        } else {
            zeta_blend(effct_feat_num, state_propagat, th);
        }
        if (B.insert_status == 0 && !EKF_stop_flg && (cfg.dev.shard_count <= 1 || reduce_fn || peers)) {  // not armed on the device (sharded map)
            double pose[24];
            state.pose24(pose);
            LM_CK(dlt_map_incremental(dev_, pose, flg_EKF_inited ? 1 : 0, &out->n_added_ds, &out->n_added_raw));
            out->added = out->n_added_ds + out->n_added_raw;
        }
        out->t_insert = wall() - t0;
        if (want_eig) {
            LM_CK(dlt_degeneracy(dev_, out->eigvals, out->eigvecs));
            out->degenerate = (out->eigvals[0] < cfg.degeneracy_eig_threshold) ? 1 : 0;
        }
    } else if (featsFromMapNum >= 5) {  // :804, one host round trip per iteration
        out->did_update = 1;
        t0 = wall();
        int rematch_num = 0;
        bool rematch_en = false, flg_EKF_converged = false;
        double K1c[kDim * 12];  // K_1[:, :12] of the last Kalman update
        double HtH12[144];
        bool have_gain = false;
        dlt_measure_out m;
        // (state.cov / LASER_POINT_COV).inverse() is the same in every iteration of one scan: the
        // covariance only changes at :1085, after the loop
        double Pinv[kDim * kDim];
        {
            double Pn[kDim * kDim];
            for (int i = 0; i < kDim * kDim; i++) Pn[i] = state.cov[i] / LASER_POINT_COV;
            if (!invert(Pn, kDim, Pinv)) {
                err = "covariance is singular";
                return DLT_E_STATE;
            }
        }
        for (int iterCount = 0; iterCount < NUM_MAX_ITERATIONS; iterCount++) {  // :820
            dlt_lio_iter rec;
            std::memset(&rec, 0, sizeof(rec));
            rec.iter = iterCount;
            rec.did_match = (iterCount == 0 || rematch_en) ? 1 : 0;  // :847
            state.pose24(rec.pose_in);
            if (reduce_fn && !peers) {  // sharded map: partial sums of this rank -> all-reduce -> host
                LM_CK(dlt_measure_dev(dev_, rec.pose_in, rec.did_match, reduce_buf_dev));
                if (reduce_fn(reduce_ctx, reduce_buf_dev, 158) != 0) {
                    err = "reduce callback failed";
                    return DLT_E_STATE;
                }
                LM_CK(dlt_fetch_result(dev_, reduce_buf_dev, &m));
            } else {
                LM_CK(dlt_measure(dev_, rec.pose_in, rec.did_match, &m));  // :829-979 on the device
            }
            effct_feat_num = m.effct_feat_num;
            const double total_residual = m.total_residual;
            rec.n_down = m.n_down;
            out->n_down = m.n_down;  // (feats_down_size arrives with the first result block)
            rec.effct_feat_num = effct_feat_num;
            rec.total_residual = total_residual;
            rec.res_mean_last = total_residual / effct_feat_num;  // :932 (nan when 0, as in the reference)
            std::memcpy(rec.HtH, m.HtH, sizeof(rec.HtH));
            std::memcpy(rec.Htr, m.Htr, sizeof(rec.Htr));

            effct_q.push_back(effct_feat_num);  // :899-918
            if ((int)effct_q.size() > QUEUE_SIZE) effct_q.pop_front();
            EKF_stop_flg = false;
            for (int v : effct_q)
                if (v <= dynamic_effect_featurepoints_threshold) {
                    EKF_stop_flg = true;
                    break;
                }
            rec.ekf_stop = EKF_stop_flg ? 1 : 0;

            // (the `!flg_EKF_inited && !EKF_stop_flg` branch at :986-1011 cannot be reached: INIT_TIME is 0
            //  and flg_EKF_inited is only cleared at :1062, right before the loop breaks at :1095-1100)
            if (!EKF_stop_flg) {  // :1012-1053
                double A[kDim * kDim];
                std::memcpy(A, Pinv, sizeof(A));
                for (int a = 0; a < 12; a++)
                    for (int b = 0; b < 12; b++) A[a * kDim + b] += m.HtH[a * 12 + b];
                // only the first 12 columns of K_1 = A^-1 are ever used (:1019)
                if (!solve_first_columns(A, kDim, 12, K1c)) {
                    err = "H^T H + (P/R)^-1 is singular";
                    return DLT_E_STATE;
                }
                std::memcpy(HtH12, m.HtH, sizeof(HtH12));
                have_gain = true;
                double vec[kDim];
                state_propagat.boxminus(state, vec);  // :1028
                // solution = K z + vec - K H vec_12 = K_1[:, :12] (H^T r - H^T H vec_12) + vec      :1032
                double rhs[12];
                for (int a = 0; a < 12; a++) {
                    double s = 0;
                    for (int b = 0; b < 12; b++) s += m.HtH[a * 12 + b] * vec[b];
                    rhs[a] = m.Htr[a] - s;
                }
                double solution[kDim];
                for (int i = 0; i < kDim; i++) {
                    double s = 0;
                    for (int a = 0; a < 12; a++) s += K1c[i * 12 + a] * rhs[a];
                    solution[i] = s + vec[i];
                }
                state.boxplus_inplace(solution);  // :1033
                std::memcpy(rec.solution, solution, sizeof(solution));
                const double rn = std::sqrt(solution[0] * solution[0] + solution[1] * solution[1] + solution[2] * solution[2]);
                const double tn = std::sqrt(solution[3] * solution[3] + solution[4] * solution[4] + solution[5] * solution[5]);
                flg_EKF_converged = (rn * 57.3 < 0.01) && (tn * 100 < 0.015);  // :1040
                last_nodegared_state = state;                                  // :1050
            } else {  // :1054-1063
                StatesGroup d = odom_to_state(th->delta_pos, th->delta_quat, th->delta_vel, th->cov_slots);
                state = last_nodegared_state.compose(d);
                flg_EKF_inited = false;
            }
            rec.converged = flg_EKF_converged ? 1 : 0;
            {
                double f[36 + kDim * kDim];
                state.to_flat(f);
                std::memcpy(rec.state_out, f, sizeof(rec.state_out));
            }
            iters.push_back(rec);

            rematch_en = false;  // :1069-1076
            if (flg_EKF_converged || ((rematch_num == 0) && (iterCount == (NUM_MAX_ITERATIONS - 2)))) {
                rematch_en = true;
                rematch_num++;
            }
            if (rematch_num >= 2 || (iterCount == NUM_MAX_ITERATIONS - 1)) {  // :1078-1094
                // The reference's covariance update (:1084-1085) is DEAD: the zeta blend that always follows (:1119 / :1127) builds
                // the new state with StatesGroup::operator+, which returns `this->cov` of last_state (common_lib.h:126, 142), and
                // nothing reads state.cov in between.  It used to cost ~5 us of host time on the critical path (the GPU idle
                // behind it); -DDLT_HOST_DEAD_COV_UPDATE=1 compiles it back in.  (k_iekf_step still evaluates it on the device.)
#ifndef DLT_HOST_DEAD_COV_UPDATE
#define DLT_HOST_DEAD_COV_UPDATE 0
#endif
                if (DLT_HOST_DEAD_COV_UPDATE && flg_EKF_inited && have_gain) {
                    // G[:, :12] = K H = K_1[:, :12] H^T H;  cov = (I - G) cov      :1084-1085
                    double G[kDim * 12], ncov[kDim * kDim];
                    for (int i = 0; i < kDim; i++)
                        for (int j = 0; j < 12; j++) {
                            double s = 0;
                            for (int a = 0; a < 12; a++) s += K1c[i * 12 + a] * HtH12[a * 12 + j];
                            G[i * 12 + j] = s;
                        }
                    for (int i = 0; i < kDim; i++)
                        for (int j = 0; j < kDim; j++) {
                            double s = state.cov[i * kDim + j];
                            for (int a = 0; a < 12; a++) s -= G[i * 12 + a] * state.cov[a * kDim + j];
                            ncov[i * kDim + j] = s;
                        }
                    std::memcpy(state.cov, ncov, sizeof(ncov));
                }
                break;
            } else if (EKF_stop_flg) {  // :1095-1101
                break;
            }
        }
        out->t_iterate = wall() - t0;
        // degradation output: eigen-decomposition of the last H^T H pose block, queued on the device now and
        // collected after map_incremental (it overlaps the host's zeta blend and the insert kernels)
        const bool want_eig = !iters.empty();
        if (want_eig) LM_CK(dlt_degeneracy_begin(dev_));

        zeta_blend(effct_feat_num, state_propagat, th);

        // ---- map_incremental(), :1164-1168
        t0 = wall();
        if (!EKF_stop_flg && (cfg.dev.shard_count <= 1 || reduce_fn || peers)) {
            double pose[24];
            state.pose24(pose);
            if (cfg.async_insert && (cfg.dev.shard_count <= 1 || peers)) {
                // off the critical path: the insert kernels run on their own stream while this call returns and the next
                // scan's deskew / VoxelGrid are enqueued; the counts are adopted by dlt_lio_collect_insert or the next scan
                LM_CK(dlt_map_incremental_async(dev_, pose, flg_EKF_inited ? 1 : 0));
                out->n_added_ds = out->n_added_raw = out->added = -1;
                insert_pending = true;
            } else {
                LM_CK(dlt_map_incremental(dev_, pose, flg_EKF_inited ? 1 : 0, &out->n_added_ds, &out->n_added_raw));
                out->added = out->n_added_ds + out->n_added_raw;
            }
        }
        out->t_insert = wall() - t0;
        if (want_eig) {
            LM_CK(dlt_degeneracy(dev_, out->eigvals, out->eigvecs));
            out->degenerate = (out->eigvals[0] < cfg.degeneracy_eig_threshold) ? 1 : 0;
        }
    }
    out->ekf_stop = EKF_stop_flg ? 1 : 0;
    out->n_iters = (int)iters.size();
    out->t_total = wall() - t_begin;
    return 0;
}

}  // namespace dlt_host

// ------------------------------------------------------------------ C ABI
using dlt_host::ImuSample;
using dlt_host::LaserMapping;

extern "C" {

void dlt_lio_default_config(dlt_lio_config *c) {
    dlt_default_config(&c->dev);
    c->max_iteration = 4;     // BASELINE configs (feat.yaml:46 ships 10)
    c->cube_len = 1000.0;     // feat.yaml:49
    c->featptsThreshold = 30; // feat.yaml:7
    c->beta = 0.1;            // feat.yaml:8
    c->det_range = 300.0f;    // laserMapping.cpp:304
    for (int i = 0; i < 3; i++) c->extrinT[i] = 0.0;
    for (int i = 0; i < 9; i++) c->extrinR[i] = (i % 4 == 0) ? 1.0 : 0.0;
    c->degeneracy_eig_threshold = 100.0;
    c->device_loop = -1;  // by measurement: host loop on one GPU, device-resident loop on a sharded map
    c->async_insert = 1;  // map_incremental off the critical path (its counts arrive with dlt_lio_collect_insert / the next scan)
}

int dlt_lio_create(const dlt_lio_config *cfg, dlt_lio *out) {
    if (!cfg || !out) return DLT_E_INVALID;
    *out = nullptr;
    LaserMapping *lm = new LaserMapping(*cfg);
    if (!lm->ok()) {
        int rc = lm->create_rc();
        delete lm;
        return rc ? rc : DLT_E_CUDA;
    }
    dlt_lio h = new dlt_lio_s;
    h->lm = lm;
    *out = h;
    return DLT_OK;
}
int dlt_lio_destroy(dlt_lio h) {
    if (!h) return DLT_E_INVALID;
    delete h->lm;
    delete h;
    return DLT_OK;
}
const char *dlt_lio_last_error(dlt_lio h) { return h ? h->lm->err.c_str() : "null handle"; }
dlt_handle dlt_lio_device(dlt_lio h) { return h ? h->lm->dev_ : nullptr; }
int dlt_lio_on_lidar_msg(dlt_lio h) {
    if (!h) return DLT_E_INVALID;
    h->lm->on_lidar_msg();
    return DLT_OK;
}
int dlt_lio_on_edge_count(dlt_lio h, int n) {
    if (!h) return DLT_E_INVALID;
    h->lm->on_edge_count(n);
    return DLT_OK;
}
int dlt_lio_force_imu_ready(dlt_lio h, const double *mean_acc3, const double *last_imu7) {
    if (!h || !mean_acc3 || !last_imu7) return DLT_E_INVALID;
    ImuSample s;
    s.t = last_imu7[0];
    std::memcpy(s.acc, last_imu7 + 1, 24);
    std::memcpy(s.gyr, last_imu7 + 4, 24);
    h->lm->imu_.force_ready(dlt_host::Vec3(mean_acc3), s);
    return DLT_OK;
}
int dlt_lio_get_state(dlt_lio h, double *s) {
    if (!h || !s) return DLT_E_INVALID;
    h->lm->state.to_flat(s);
    return DLT_OK;
}
int dlt_lio_set_state(dlt_lio h, const double *s, int also_last) {
    if (!h || !s) return DLT_E_INVALID;
    h->lm->state.from_flat(s);
    if (also_last) {
        h->lm->last_state.from_flat(s);
        h->lm->last_nodegared_state.from_flat(s);
    }
    return DLT_OK;
}
int dlt_lio_get_flags(dlt_lio h, int *f) {
    if (!h || !f) return DLT_E_INVALID;
    LaserMapping *l = h->lm;
    f[0] = l->EKF_stop_flg;
    f[1] = l->flg_EKF_inited;
    f[2] = l->dynamic_effect_featurepoints_threshold;
    f[3] = (int)l->lidar_cnt;
    f[4] = l->Localmap_Initialized;
    f[5] = (int)l->effct_q.size();
    f[6] = l->map_built;
    f[7] = l->imu_.need_init() ? 0 : 1;
    return DLT_OK;
}
int dlt_lio_get_localmap(dlt_lio h, float *b) {
    if (!h || !b) return DLT_E_INVALID;
    std::memcpy(b, h->lm->LocalMap_min, 12);
    std::memcpy(b + 3, h->lm->LocalMap_max, 12);
    return DLT_OK;
}
int dlt_lio_process_scan(dlt_lio h, const void *pts48, int n, double lidar_beg_time, const double *imu7, int n_imu,
                         const dlt_lio_thermal *thermal, dlt_lio_scan_out *out) {
    if (!h || !out || n < 0 || n_imu < 0 || (n > 0 && !pts48) || (n_imu > 0 && !imu7)) return DLT_E_INVALID;
    static_assert(sizeof(ImuSample) == 7 * sizeof(double), "imu7 layout");
    return h->lm->process_scan(pts48, n, lidar_beg_time, reinterpret_cast<const ImuSample *>(imu7), n_imu, thermal, out);
}
int dlt_lio_prefetch_scan(dlt_lio h, const void *pts48, int n) {
    if (!h) return DLT_E_INVALID;
    int rc = dlt_scan_prefetch(h->lm->dev_, pts48, n);
    if (rc != 0) h->lm->err = dlt_last_error(h->lm->dev_);
    return rc;
}
int dlt_lio_process_scan_dev(dlt_lio h, const void *pts48_dev, int n, double lidar_beg_time, double observation_end_time, const double *imu7,
                             int n_imu, const dlt_lio_thermal *thermal, dlt_lio_scan_out *out) {
    if (!h || !out || n < 0 || n_imu < 0 || (n > 0 && !pts48_dev) || (n_imu > 0 && !imu7)) return DLT_E_INVALID;
    return h->lm->process_scan(pts48_dev, n, lidar_beg_time, reinterpret_cast<const ImuSample *>(imu7), n_imu, thermal, out, true,
                               observation_end_time);
}
int dlt_lio_process_cloud(dlt_lio h, const void *cloud_data, int n_points, const dlt_cloud_layout *layout, int sensor, int point_filter_num,
                          float lidar_min_range, float lidar_max_range, double header_stamp, const double *imu7, int n_imu,
                          const dlt_lio_thermal *thermal, dlt_lio_scan_out *out, int *n_sampled) {
    if (!h || !out || !layout || n_points < 0 || n_imu < 0 || (n_points > 0 && !cloud_data) || (n_imu > 0 && !imu7)) return DLT_E_INVALID;
    void *pts_dev = nullptr;
    int n = 0;
    double timespan = 0, sweep = 0, shift = 0;
    int rc = dlt_frontend_sample(h->lm->dev_, cloud_data, n_points, layout, sensor, point_filter_num, lidar_min_range, lidar_max_range, &pts_dev, &n,
                                 &timespan, &sweep, &shift);
    if (rc != 0) {
        h->lm->err = dlt_last_error(h->lm->dev_);
        return rc;
    }
    if (n_sampled) *n_sampled = n;
    const double lidar_beg_time = header_stamp - shift;  // feature_extract.cpp:383 (RoboSense), else the header stamp
    // observation_end_time = lidar_beg_time + points.back().normal_z, laserMapping.cpp:546
    return h->lm->process_scan(pts_dev, n, lidar_beg_time, reinterpret_cast<const ImuSample *>(imu7), n_imu, thermal, out, true, lidar_beg_time + sweep);
}
int dlt_lio_set_reduce(dlt_lio h, dlt_lio_reduce_fn reduce, void *ctx, double *result_dev) {
    if (!h) return DLT_E_INVALID;
    if (reduce && !result_dev) result_dev = dlt_result_dev(h->lm->dev_);  // the handle's own buffer
    h->lm->reduce_fn = reduce;
    h->lm->reduce_ctx = ctx;
    h->lm->reduce_buf_dev = result_dev;
    // the same sum over the ranks also carries the per-point map_incremental decisions of a sharded map
    return dlt_set_shard_reduce(h->lm->dev_, reduce ? &dlt_host::LaserMapping::reduce_trampoline : nullptr, h->lm);
}
int dlt_lio_peer_export(dlt_lio h, unsigned char *blob) {
    if (!h) return DLT_E_INVALID;
    int rc = dlt_peer_export(h->lm->dev_, blob);
    if (rc) h->lm->err = dlt_last_error(h->lm->dev_);
    return rc;
}
int dlt_lio_peer_attach(dlt_lio h, const unsigned char *blobs) {
    if (!h) return DLT_E_INVALID;
    int rc = dlt_peer_attach(h->lm->dev_, blobs);
    if (rc) h->lm->err = dlt_last_error(h->lm->dev_);
    h->lm->peers = rc == 0;
    return rc;
}
int dlt_lio_peer_detach(dlt_lio h) {
    if (!h) return DLT_E_INVALID;
    h->lm->peers = false;
    return dlt_peer_detach(h->lm->dev_);
}
int dlt_lio_collect_insert(dlt_lio h, int *n_added_ds, int *n_added_raw) {
    if (!h) return DLT_E_INVALID;
    return h->lm->collect_insert(n_added_ds, n_added_raw);
}
int dlt_lio_get_iters(dlt_lio h, dlt_lio_iter *iters, int cap) {
    if (!h) return DLT_E_INVALID;
    int n = (int)h->lm->iters.size();
    for (int i = 0; i < n && i < cap; i++) iters[i] = h->lm->iters[i];
    return n;
}
int dlt_lio_get_imu_poses(dlt_lio h, double *pose22, int cap) {
    if (!h) return DLT_E_INVALID;
    int n = (int)h->lm->imu_.IMUpose.size();
    for (int i = 0; i < n && i < cap; i++) std::memcpy(pose22 + 22 * (size_t)i, &h->lm->imu_.IMUpose[i], sizeof(dlt_host::Pose6D));
    return n;
}

}  // extern "C"
