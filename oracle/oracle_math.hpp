// oracle/oracle_math.hpp -- TEST INFRASTRUCTURE ONLY (CPU parity oracle, kind "port").
//
// Plain C++ restatement of the small-matrix / SO(3) / state algebra the reference's
// eskf_lio hot path uses through Eigen.  Eigen is not installed in this image, so the
// arithmetic is written out by hand; every function cites the reference lines it
// follows (paths relative to /root/reference/).  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may use anything under oracle/.
//
// Parity status: the fp64 algebra here follows the reference's formulas; Eigen's
// internal summation order is not reproduced (unverifiable without Eigen), which is
// why fp64 outputs are compared with a relative tolerance, never bit-for-bit.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace orc {

static const int DIM = 24;  // DIM_OF_STATES, common_lib.h:24

struct V3 {
    double v[3];
    V3() { v[0] = v[1] = v[2] = 0.0; }
    V3(double a, double b, double c) { v[0] = a; v[1] = b; v[2] = c; }
    double &operator[](int i) { return v[i]; }
    const double &operator[](int i) const { return v[i]; }
};
inline V3 operator+(const V3 &a, const V3 &b) { return V3(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline V3 operator-(const V3 &a, const V3 &b) { return V3(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
inline V3 operator*(const V3 &a, double s) { return V3(a[0] * s, a[1] * s, a[2] * s); }
inline V3 operator*(double s, const V3 &a) { return a * s; }
inline V3 operator/(const V3 &a, double s) { return V3(a[0] / s, a[1] / s, a[2] / s); }
inline double dot(const V3 &a, const V3 &b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline double norm(const V3 &a) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

// row-major 3x3
struct M3 {
    double m[9];
    M3() { std::memset(m, 0, sizeof(m)); }
    static M3 I() {
        M3 r;
        r.m[0] = r.m[4] = r.m[8] = 1.0;
        return r;
    }
    double &operator()(int i, int j) { return m[3 * i + j]; }
    const double &operator()(int i, int j) const { return m[3 * i + j]; }
};
inline M3 operator*(const M3 &a, const M3 &b) {
    M3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r(i, j) = a(i, 0) * b(0, j) + a(i, 1) * b(1, j) + a(i, 2) * b(2, j);
    return r;
}
inline V3 operator*(const M3 &a, const V3 &b) {
    V3 r;
    for (int i = 0; i < 3; i++) r[i] = a(i, 0) * b[0] + a(i, 1) * b[1] + a(i, 2) * b[2];
    return r;
}
inline M3 operator*(const M3 &a, double s) {
    M3 r;
    for (int i = 0; i < 9; i++) r.m[i] = a.m[i] * s;
    return r;
}
inline M3 operator+(const M3 &a, const M3 &b) {
    M3 r;
    for (int i = 0; i < 9; i++) r.m[i] = a.m[i] + b.m[i];
    return r;
}
inline M3 transpose(const M3 &a) {
    M3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r(i, j) = a(j, i);
    return r;
}
// SKEW_SYM_MATRX, so3_math.h:9
inline M3 skew(const V3 &v) {
    M3 k;
    k(0, 1) = -v[2];
    k(0, 2) = v[1];
    k(1, 0) = v[2];
    k(1, 2) = -v[0];
    k(2, 0) = -v[1];
    k(2, 1) = v[0];
    return k;
}
inline double trace(const M3 &a) { return a.m[0] + a.m[4] + a.m[8]; }

// Exp(ang_vel, dt): so3_math.h:32-52 (threshold 1e-7 on |ang_vel|)
inline M3 Exp(const V3 &ang_vel, double dt) {
    double n = norm(ang_vel);
    if (n > 0.0000001) {
        V3 axis = ang_vel / n;
        M3 K = skew(axis);
        double r_ang = n * dt;
        return M3::I() + K * std::sin(r_ang) + (K * (1.0 - std::cos(r_ang))) * K;
    }
    return M3::I();
}
// Exp(v1,v2,v3): so3_math.h:54-72 (threshold 1e-5 on the norm)
inline M3 Exp3(double v1, double v2, double v3) {
    double n = std::sqrt(v1 * v1 + v2 * v2 + v3 * v3);
    if (n > 0.00001) {
        V3 axis(v1 / n, v2 / n, v3 / n);
        M3 K = skew(axis);
        return M3::I() + K * std::sin(n) + (K * (1.0 - std::cos(n))) * K;
    }
    return M3::I();
}
// Log(R): so3_math.h:75-81
inline V3 Log(const M3 &R) {
    double tr = trace(R);
    double theta = (tr > 3.0 - 1e-6) ? 0.0 : std::acos(0.5 * (tr - 1));
    V3 K(R(2, 1) - R(1, 2), R(0, 2) - R(2, 0), R(1, 0) - R(0, 1));
    return (std::abs(theta) < 0.001) ? (K * 0.5) : (K * (0.5 * theta / std::sin(theta)));
}
// RotMtoEuler: so3_math.h:83-103
inline V3 RotMtoEuler(const M3 &rot) {
    double sy = std::sqrt(rot(0, 0) * rot(0, 0) + rot(1, 0) * rot(1, 0));
    bool singular = sy < 1e-6;
    double x, y, z;
    if (!singular) {
        x = std::atan2(rot(2, 1), rot(2, 2));
        y = std::atan2(-rot(2, 0), sy);
        z = std::atan2(rot(1, 0), rot(0, 0));
    } else {
        x = std::atan2(-rot(1, 2), rot(1, 1));
        y = std::atan2(-rot(2, 0), sy);
        z = 0;
    }
    return V3(x, y, z);
}
// Eigen::Quaterniond(w,x,y,z).toRotationMatrix() (used by odomToStateGruop, laserMapping.cpp:225-226)
inline M3 quat_to_rot(double w, double x, double y, double z) {
    M3 r;
    double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    double twx = tx * w, twy = ty * w, twz = tz * w;
    double txx = tx * x, txy = ty * x, txz = tz * x;
    double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    r(0, 0) = 1 - (tyy + tzz);
    r(0, 1) = txy - twz;
    r(0, 2) = txz + twy;
    r(1, 0) = txy + twz;
    r(1, 1) = 1 - (txx + tzz);
    r(1, 2) = tyz - twx;
    r(2, 0) = txz - twy;
    r(2, 1) = tyz + twx;
    r(2, 2) = 1 - (txx + tyy);
    return r;
}

// ---------------------------------------------------------------- dense NxN helpers (row-major)
typedef std::vector<double> Mat;  // n*m row-major

inline Mat mat_mul(const Mat &A, int n, int k, const Mat &B, int m) {  // (n x k) * (k x m)
    Mat C((size_t)n * m, 0.0);
    for (int i = 0; i < n; i++)
        for (int l = 0; l < k; l++) {
            double a = A[(size_t)i * k + l];
            if (a == 0.0) continue;
            for (int j = 0; j < m; j++) C[(size_t)i * m + j] += a * B[(size_t)l * m + j];
        }
    return C;
}
inline Mat mat_T(const Mat &A, int n, int m) {
    Mat T((size_t)n * m);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < m; j++) T[(size_t)j * n + i] = A[(size_t)i * m + j];
    return T;
}
// General inverse by LU with partial pivoting (Eigen's .inverse() for n>4 is PartialPivLU).
inline bool mat_inverse(const Mat &Ain, int n, Mat &inv) {
    Mat A(Ain);
    inv.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++) inv[(size_t)i * n + i] = 1.0;
    for (int c = 0; c < n; c++) {
        int piv = c;
        double best = std::fabs(A[(size_t)c * n + c]);
        for (int r = c + 1; r < n; r++) {
            double v = std::fabs(A[(size_t)r * n + c]);
            if (v > best) {
                best = v;
                piv = r;
            }
        }
        if (best == 0.0) return false;
        if (piv != c) {
            for (int j = 0; j < n; j++) {
                std::swap(A[(size_t)c * n + j], A[(size_t)piv * n + j]);
                std::swap(inv[(size_t)c * n + j], inv[(size_t)piv * n + j]);
            }
        }
        double d = A[(size_t)c * n + c];
        for (int r = c + 1; r < n; r++) {
            double f = A[(size_t)r * n + c] / d;
            if (f == 0.0) continue;
            for (int j = c; j < n; j++) A[(size_t)r * n + j] -= f * A[(size_t)c * n + j];
            for (int j = 0; j < n; j++) inv[(size_t)r * n + j] -= f * inv[(size_t)c * n + j];
        }
    }
    for (int c = n - 1; c >= 0; c--) {
        double d = A[(size_t)c * n + c];
        for (int j = 0; j < n; j++) inv[(size_t)c * n + j] /= d;
        for (int r = 0; r < c; r++) {
            double f = A[(size_t)r * n + c];
            if (f == 0.0) continue;
            for (int j = 0; j < n; j++) inv[(size_t)r * n + j] -= f * inv[(size_t)c * n + j];
        }
    }
    return true;
}

// ---------------------------------------------------------------- StatesGroup, common_lib.h:73-228
struct State {
    M3 rot_end;
    V3 pos_end;
    M3 R_L_I;
    V3 T_L_I;
    V3 vel_end, bias_g, bias_a, gravity;
    double cov[DIM * DIM];
    State() {  // common_lib.h:75-86, INIT_COV = 1
        rot_end = M3::I();
        R_L_I = M3::I();
        std::memset(cov, 0, sizeof(cov));
        for (int i = 0; i < DIM; i++) cov[i * DIM + i] = 1.0;
    }
};
// operator+(vector), common_lib.h:115-129 (result carries this->cov)
inline State state_plus(const State &s, const double *add) {
    State a;
    a.rot_end = s.rot_end * Exp3(add[0], add[1], add[2]);
    a.pos_end = s.pos_end + V3(add[3], add[4], add[5]);
    a.R_L_I = s.R_L_I * Exp3(add[6], add[7], add[8]);
    a.T_L_I = s.T_L_I + V3(add[9], add[10], add[11]);
    a.vel_end = s.vel_end + V3(add[12], add[13], add[14]);
    a.bias_g = s.bias_g + V3(add[15], add[16], add[17]);
    a.bias_a = s.bias_a + V3(add[18], add[19], add[20]);
    a.gravity = s.gravity + V3(add[21], add[22], add[23]);
    std::memcpy(a.cov, s.cov, sizeof(a.cov));
    return a;
}
// operator+(StatesGroup), common_lib.h:131-144
inline State state_plus_state(const State &s, const State &b) {
    State r;
    r.rot_end = s.rot_end * b.rot_end;
    r.pos_end = s.pos_end + b.pos_end;
    r.R_L_I = s.R_L_I * b.R_L_I;
    r.T_L_I = s.T_L_I + b.T_L_I;
    r.vel_end = s.vel_end + b.vel_end;
    r.bias_g = s.bias_g;
    r.bias_a = s.bias_a;
    r.gravity = s.gravity;
    std::memcpy(r.cov, s.cov, sizeof(r.cov));
    return r;
}
// operator+=(vector), common_lib.h:147-158 (cov untouched)
inline void state_add_inplace(State &s, const double *add) {
    s.rot_end = s.rot_end * Exp3(add[0], add[1], add[2]);
    s.pos_end = s.pos_end + V3(add[3], add[4], add[5]);
    s.R_L_I = s.R_L_I * Exp3(add[6], add[7], add[8]);
    s.T_L_I = s.T_L_I + V3(add[9], add[10], add[11]);
    s.vel_end = s.vel_end + V3(add[12], add[13], add[14]);
    s.bias_g = s.bias_g + V3(add[15], add[16], add[17]);
    s.bias_a = s.bias_a + V3(add[18], add[19], add[20]);
    s.gravity = s.gravity + V3(add[21], add[22], add[23]);
}
// operator-(StatesGroup), common_lib.h:173-187:  a = this (-) b
inline void state_minus(const State &s, const State &b, double *a) {
    M3 rotd = transpose(b.rot_end) * s.rot_end;
    M3 rotLI = transpose(b.R_L_I) * s.R_L_I;
    V3 l0 = Log(rotd), l1 = Log(rotLI);
    for (int i = 0; i < 3; i++) {
        a[i] = l0[i];
        a[3 + i] = s.pos_end[i] - b.pos_end[i];
        a[6 + i] = l1[i];
        a[9 + i] = s.T_L_I[i] - b.T_L_I[i];
        a[12 + i] = s.vel_end[i] - b.vel_end[i];
        a[15 + i] = s.bias_g[i] - b.bias_g[i];
        a[18 + i] = s.bias_a[i] - b.bias_a[i];
        a[21 + i] = s.gravity[i] - b.gravity[i];
    }
}
// operator*(scale), common_lib.h:190-205
inline State state_scale(const State &s, double scale) {
    State a;
    V3 so3 = Log(s.rot_end);
    a.rot_end = Exp3(so3[0] * scale, so3[1] * scale, so3[2] * scale);
    a.pos_end = s.pos_end * scale;
    a.R_L_I = s.R_L_I;
    a.T_L_I = s.T_L_I * scale;
    a.vel_end = s.vel_end * scale;
    a.bias_g = s.bias_g * scale;
    a.bias_a = s.bias_a * scale;
    a.gravity = s.gravity * scale;
    std::memcpy(a.cov, s.cov, sizeof(a.cov));
    return a;
}

// flat (de)serialisation used by the C API: rot_end[9] pos_end[3] R_L_I[9] T_L_I[3]
// vel_end[3] bias_g[3] bias_a[3] gravity[3] = 36 doubles, then cov[576].
inline void state_to_flat(const State &s, double *f) {
    std::memcpy(f, s.rot_end.m, 72);
    std::memcpy(f + 9, s.pos_end.v, 24);
    std::memcpy(f + 12, s.R_L_I.m, 72);
    std::memcpy(f + 21, s.T_L_I.v, 24);
    std::memcpy(f + 24, s.vel_end.v, 24);
    std::memcpy(f + 27, s.bias_g.v, 24);
    std::memcpy(f + 30, s.bias_a.v, 24);
    std::memcpy(f + 33, s.gravity.v, 24);
    std::memcpy(f + 36, s.cov, sizeof(s.cov));
}
inline void state_from_flat(State &s, const double *f) {
    std::memcpy(s.rot_end.m, f, 72);
    std::memcpy(s.pos_end.v, f + 9, 24);
    std::memcpy(s.R_L_I.m, f + 12, 72);
    std::memcpy(s.T_L_I.v, f + 21, 24);
    std::memcpy(s.vel_end.v, f + 24, 24);
    std::memcpy(s.bias_g.v, f + 27, 24);
    std::memcpy(s.bias_a.v, f + 30, 24);
    std::memcpy(s.gravity.v, f + 33, 24);
    std::memcpy(s.cov, f + 36, sizeof(s.cov));
}

// ---------------------------------------------------------------- symmetric eigen (cyclic Jacobi)
// NEW output (no reference counterpart): eigen-decomposition of the 6x6
// pose block of H^T H.  Ascending eigenvalues; eigenvectors in the columns of V (row-major).
// "parity unpinned": checked by properties only (reconstruction, orthogonality, trace).
inline void jacobi_eig_sym(const double *Ain, int n, double *evals, double *V) {
    std::vector<double> A(Ain, Ain + n * n);
    for (int i = 0; i < n * n; i++) V[i] = 0.0;
    for (int i = 0; i < n; i++) V[i * n + i] = 1.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) (i == j ? diag : off) += A[i * n + j] * A[i * n + j];
        if (off <= 1e-30 * diag || off == 0.0) break;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                double apq = A[p * n + q];
                if (apq == 0.0) continue;
                double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; k++) {
                    double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - s * akq;
                    A[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {
                    double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - s * aqk;
                    A[q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; k++) {
                    double vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = c * vkp - s * vkq;
                    V[k * n + q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < n; i++) evals[i] = A[i * n + i];
    // selection sort ascending, permuting eigenvector columns
    for (int i = 0; i < n - 1; i++) {
        int m = i;
        for (int j = i + 1; j < n; j++)
            if (evals[j] < evals[m]) m = j;
        if (m != i) {
            std::swap(evals[i], evals[m]);
            for (int k = 0; k < n; k++) std::swap(V[k * n + i], V[k * n + m]);
        }
    }
}

}  // namespace orc
