// oracle/oracle.hpp -- TEST INFRASTRUCTURE ONLY (CPU parity oracle, kind "port").
//
// Plain C++ restatement of DaLiTI's eskf_lio scan-to-map hot path, one function per
// reference function, each citing the reference lines it follows (paths relative to
// /root/reference/).  What can be compiled from the reference itself (the ikd-Tree) is
// NOT restated here as the primary checker: oracle/_ref/libikd_ref.so is the unmodified
// reference code and `MapBackend` below can be backed by it; `PortMap` is an
// independent restatement of the ikd-Tree's *contents and query semantics* that is
// itself pinned against the reference library by tests/test_oracle_map.py.
//
// PARITY STATUS (also stated in DESIGN.md):
//   * ikd-Tree semantics (kNN, downsample-on-insert, box delete): PINNED against the
//     unmodified reference code run here (oracle/_ref).
//   * esti_plane: follows Eigen 3.3.7 ColPivHouseholderQR step by step, but Eigen is not
//     installed and the reference ships no golden vectors  ->  "parity unpinned" vs real
//     Eigen; property-checked (fp64 SVD plane agreement, residual gate).
//   * pcl::VoxelGrid (PCL 1.10, not vendored in the reference): "parity unpinned" vs real
//     PCL; voxel assignment is integer-exact by construction.
//   * UndistortPcl / IEKF loop / Kalman algebra / map_incremental / fov segment: follow
//     the reference source line by line; the reference ships no tests or golden files
//      -> "parity unpinned" beyond this restatement.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may include, link or execute anything under oracle/.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <memory>
#include <queue>
#include <unordered_map>
#include <vector>

#include "oracle_math.hpp"

namespace orc {

// pcl::PointXYZINormal (my_utility.h:57): 48 bytes,
//   x y z 1 | normal_x(time ratio) normal_y(ring) normal_z(span s) 0 | intensity curvature pad pad
// (feature_extract.cpp:337-346 packs the time fields).
struct Pt {
    float x, y, z, d3;
    float nx, ny, nz, dn3;
    float intensity, curvature, p0, p1;
};
static_assert(sizeof(Pt) == 48, "PointXYZINormal layout");

struct P4 {  // x y z intensity
    float x, y, z, w;
};

// calc_dist: common_lib.h:244-248 and ikd_Tree.cpp:1682-1688 (same float expression,
// x86-64 -O3 without -mfma: no contraction)
inline float calc_dist(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = ax - bx, dy = ay - by, dz = az - bz;
    float d = dx * dx + dy * dy;
    d = d + dz * dz;
    return d;
}

// ======================================================================== esti_plane
// common_lib.h:267-299.  A (5x3, float) n = -1 solved with
// Eigen::ColPivHouseholderQR<Matrix<float,5,3>> (Eigen 3.3.7: ColPivHouseholderQR.h
// computeInPlace / _solve_impl, Householder.h makeHouseholder / applyHouseholderOnTheLeft).
// Reductions follow Eigen's SSE2 order where it is determined by the expression type:
//   size-5 / size-4 column norms: ((a0+a2)+(a1+a3)) [+a4]; dynamic size <4: left to right;
//   fixed Vector3f norm: a0+(a1+a2).  The essential^T*block GEMV is summed left to right
//   (Eigen's order there depends on run-time pointer alignment and cannot be restated).
namespace detail {
inline float sum_sq_dyn(const float *v, int n) {  // squaredNorm of a dynamic-size vector
    if (n >= 4) {
        float a0 = v[0] * v[0], a1 = v[1] * v[1], a2 = v[2] * v[2], a3 = v[3] * v[3];
        float r = (a0 + a2) + (a1 + a3);
        for (int i = 4; i < n; i++) r = r + v[i] * v[i];
        return r;
    }
    if (n <= 0) return 0.f;
    float r = v[0] * v[0];
    for (int i = 1; i < n; i++) r = r + v[i] * v[i];
    return r;
}
}  // namespace detail

inline bool esti_plane(float pabcd[4], const float pts[5][3], float threshold) {
    const int rows = 5, cols = 3, size = 3;
    // column-major working copy: a[c][r]
    float a[3][5];
    for (int j = 0; j < rows; j++)
        for (int c = 0; c < cols; c++) a[c][j] = pts[j][c];
    float hc[3];
    int transp[3];
    float nu[3], nd[3];
    const float eps = FLT_EPSILON;
    for (int k = 0; k < cols; k++) {
        nd[k] = std::sqrt(detail::sum_sq_dyn(a[k], rows));
        nu[k] = nd[k];
    }
    float maxn = nu[0];
    for (int k = 1; k < cols; k++)
        if (nu[k] > maxn) maxn = nu[k];
    float th = maxn * eps;
    float threshold_helper = (th * th) / float(rows);
    float norm_downdate_threshold = std::sqrt(eps);
    int nonzero_pivots = size;
    float maxpivot = 0.f;
    for (int k = 0; k < size; k++) {
        int big = k;
        float bigv = nu[k];
        for (int j = k + 1; j < cols; j++)
            if (nu[j] > bigv) {
                bigv = nu[j];
                big = j;
            }
        float big_sq = bigv * bigv;
        if (nonzero_pivots == size && big_sq < threshold_helper * float(rows - k)) nonzero_pivots = k;
        transp[k] = big;
        if (k != big) {
            for (int r = 0; r < rows; r++) std::swap(a[k][r], a[big][r]);
            std::swap(nu[k], nu[big]);
            std::swap(nd[k], nd[big]);
        }
        // makeHouseholderInPlace on a[k][k..4]
        int tl = rows - k - 1;
        float *tail = &a[k][k + 1];
        float tailSq = (tl == 0) ? 0.f : detail::sum_sq_dyn(tail, tl);
        float c0 = a[k][k];
        float beta, tau;
        if (tailSq <= FLT_MIN) {
            tau = 0.f;
            beta = c0;
            for (int i = 0; i < tl; i++) tail[i] = 0.f;
        } else {
            beta = std::sqrt(c0 * c0 + tailSq);
            if (c0 >= 0.f) beta = -beta;
            float den = c0 - beta;
            for (int i = 0; i < tl; i++) tail[i] = tail[i] / den;
            tau = (beta - c0) / beta;
        }
        hc[k] = tau;
        a[k][k] = beta;
        if (std::fabs(beta) > maxpivot) maxpivot = std::fabs(beta);
        // apply H_k to the trailing columns (rows k..4)
        if (tau != 0.f) {
            for (int j = k + 1; j < cols; j++) {
                float tmp = 0.f;
                for (int i = 0; i < tl; i++) tmp = (i == 0) ? tail[0] * a[j][k + 1] : tmp + tail[i] * a[j][k + 1 + i];
                tmp = tmp + a[j][k];
                a[j][k] = a[j][k] - tau * tmp;
                for (int i = 0; i < tl; i++) a[j][k + 1 + i] = a[j][k + 1 + i] - tmp * (tau * tail[i]);
            }
        }
        // LAPACK-style column norm downdate
        for (int j = k + 1; j < cols; j++) {
            if (nu[j] != 0.f) {
                float temp = std::fabs(a[j][k]) / nu[j];
                temp = (1.f + temp) * (1.f - temp);
                temp = temp < 0.f ? 0.f : temp;
                float q = nu[j] / nd[j];
                float temp2 = temp * (q * q);
                if (temp2 <= norm_downdate_threshold) {
                    nd[j] = std::sqrt(detail::sum_sq_dyn(&a[j][k + 1], rows - k - 1));
                    nu[j] = nd[j];
                } else {
                    nu[j] = nu[j] * std::sqrt(temp);
                }
            }
        }
    }
    (void)maxpivot;
    int perm[3] = {0, 1, 2};
    for (int k = 0; k < size; k++) std::swap(perm[k], perm[transp[k]]);

    // solve: c = Q^T b, R x = c, undo permutation
    float nv[3] = {0.f, 0.f, 0.f};
    if (nonzero_pivots > 0) {
        float c[5] = {-1.f, -1.f, -1.f, -1.f, -1.f};
        for (int k = 0; k < nonzero_pivots; k++) {
            int tl = rows - k - 1;
            const float *ess = &a[k][k + 1];
            float tau = hc[k];
            if (tau != 0.f) {
                float tmp = 0.f;
                for (int i = 0; i < tl; i++) tmp = (i == 0) ? ess[0] * c[k + 1] : tmp + ess[i] * c[k + 1 + i];
                tmp = tmp + c[k];
                c[k] = c[k] - tau * tmp;
                for (int i = 0; i < tl; i++) c[k + 1 + i] = c[k + 1 + i] - tmp * (tau * ess[i]);
            }
        }
        for (int i = nonzero_pivots - 1; i >= 0; i--) {
            c[i] = c[i] / a[i][i];
            for (int r = 0; r < i; r++) c[r] = c[r] - c[i] * a[i][r];
        }
        for (int i = 0; i < nonzero_pivots; i++) nv[perm[i]] = c[i];
        for (int i = nonzero_pivots; i < cols; i++) nv[perm[i]] = 0.f;
    }
    float n = std::sqrt(nv[0] * nv[0] + (nv[1] * nv[1] + nv[2] * nv[2]));
    pabcd[0] = nv[0] / n;
    pabcd[1] = nv[1] / n;
    pabcd[2] = nv[2] / n;
    pabcd[3] = (float)(1.0 / (double)n);
    for (int j = 0; j < rows; j++) {
        float v = pabcd[0] * pts[j][0] + pabcd[1] * pts[j][1];
        v = v + pabcd[2] * pts[j][2];
        v = v + pabcd[3];
        if (std::fabs(v) > threshold) return false;
    }
    return true;
}

// ======================================================================== pcl::VoxelGrid
// Call sites laserMapping.cpp:153,703,775-776.  Restates PCL 1.10
// pcl::VoxelGrid<PointT>::applyFilter (filters/impl/voxel_grid.hpp) with
// downsample_all_data_=true, min_points_per_voxel_=0, no filter field, and
// pcl::CentroidPoint's XYZ / Normal / Intensity / Curvature accumulators
// (common/impl/accumulators.hpp).  `stable` selects std::stable_sort instead of PCL's
// std::sort so that the within-voxel summation order is the ascending input index (the
// order PCL's unstable sort leaves is unspecified).
// voxel_of_point (optional, size n): output slot of every input point.
inline int voxel_grid(const Pt *in, int n, float leaf, std::vector<Pt> &out, bool stable,
                      std::vector<int> *voxel_of_point = nullptr, std::vector<unsigned> *idx_of_point = nullptr) {
    out.clear();
    if (n <= 0) return 0;
    float inv = 1.0f / leaf;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = 0; i < n; i++) {
        const float p[3] = {in[i].x, in[i].y, in[i].z};
        for (int a = 0; a < 3; a++) {
            mn[a] = std::min(mn[a], p[a]);
            mx[a] = std::max(mx[a], p[a]);
        }
    }
    int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1;
    int64_t dy = (int64_t)((mx[1] - mn[1]) * inv) + 1;
    int64_t dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
    if (dx * dy * dz > (int64_t)INT32_MAX) {  // PCL: warn and pass the input through
        out.assign(in, in + n);
        if (voxel_of_point) {
            voxel_of_point->resize(n);
            for (int i = 0; i < n; i++) (*voxel_of_point)[i] = i;
        }
        return n;
    }
    int min_b[3], max_b[3], div_b[3];
    for (int a = 0; a < 3; a++) {
        min_b[a] = (int)std::floor(mn[a] * inv);
        max_b[a] = (int)std::floor(mx[a] * inv);
        div_b[a] = max_b[a] - min_b[a] + 1;
    }
    int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
    struct CI {
        unsigned idx;
        unsigned pi;
    };
    std::vector<CI> iv(n);
    for (int i = 0; i < n; i++) {
        int i0 = (int)(std::floor(in[i].x * inv) - (float)min_b[0]);
        int i1 = (int)(std::floor(in[i].y * inv) - (float)min_b[1]);
        int i2 = (int)(std::floor(in[i].z * inv) - (float)min_b[2]);
        iv[i].idx = (unsigned)(i0 * mul[0] + i1 * mul[1] + i2 * mul[2]);
        iv[i].pi = (unsigned)i;
    }
    if (idx_of_point) {
        idx_of_point->resize(n);
        for (int i = 0; i < n; i++) (*idx_of_point)[i] = iv[i].idx;
    }
    auto less = [](const CI &a, const CI &b) { return a.idx < b.idx; };
    if (stable)
        std::stable_sort(iv.begin(), iv.end(), less);
    else
        std::sort(iv.begin(), iv.end(), less);
    if (voxel_of_point) voxel_of_point->assign(n, -1);
    size_t index = 0;
    while (index < iv.size()) {
        size_t i = index + 1;
        while (i < iv.size() && iv[i].idx == iv[index].idx) ++i;
        float sx = 0.f, sy = 0.f, sz = 0.f;              // AccumulatorXYZ
        float snx = 0.f, sny = 0.f, snz = 0.f, snw = 0.f;  // AccumulatorNormal (Vector4f)
        float si = 0.f, sc = 0.f;                          // AccumulatorIntensity / Curvature
        for (size_t li = index; li < i; li++) {
            const Pt &p = in[iv[li].pi];
            sx += p.x;
            sy += p.y;
            sz += p.z;
            snx += p.nx;
            sny += p.ny;
            snz += p.nz;
            snw += p.dn3;
            si += p.intensity;
            sc += p.curvature;
            if (voxel_of_point) (*voxel_of_point)[iv[li].pi] = (int)out.size();
        }
        float cnt = (float)(i - index);
        Pt o;
        std::memset(&o, 0, sizeof(o));
        o.x = sx / cnt;
        o.y = sy / cnt;
        o.z = sz / cnt;
        o.d3 = 1.0f;
        float nn = std::sqrt((snx * snx + snz * snz) + (sny * sny + snw * snw));
        if (nn > 0.f) {
            o.nx = snx / nn;
            o.ny = sny / nn;
            o.nz = snz / nn;
        }
        o.intensity = si / cnt;
        o.curvature = sc / cnt;
        out.push_back(o);
        index = i;
    }
    return (int)out.size();
}

// ======================================================================== IMU process / deskew
// ------------------------------------------------------------------ LiDAR front end, feature_enabled = 0
// FeatureExtract::cachePointCloud (eskf_lio/src/feature_extract.cpp:264-423) + samplePointCloud (:425-450), restated
// over raw PointCloud2 records (point_step + field offsets) with the sensor structs of eskf_lio/include/my_utility.h:19-54.
// TEST_LIO_SAM_6AXIS_DATA is #defined in the reference (:262), so the Velodyne branch is the 6-axis one.
struct CloudLayout {
    int point_step, off_x, off_y, off_z, off_intensity, off_ring, off_time;
};
enum { SENSOR_VELODYNE = 0, SENSOR_LIVOX = 1, SENSOR_OUSTER = 2, SENSOR_ROBOSENSE = 3 };

inline float pointDistance(const Pt &p) {  // my_utility.h:76-79 (sqrt of a float sum: correctly rounded either way)
    float s = p.x * p.x + p.y * p.y;
    s = s + p.z * p.z;
    return std::sqrt(s);
}

inline int frontend_sample(const unsigned char *data, int n, const CloudLayout &L, int sensor, int point_filter_num, float lidarMinRange,
                           float lidarMaxRange, std::vector<Pt> &sampleCloud, double &timespan_out, double &stamp_shift) {
    auto rec = [&](size_t i) { return data + i * (size_t)L.point_step; };
    auto f32 = [](const unsigned char *p) { float v; std::memcpy(&v, p, 4); return v; };
    auto u32 = [](const unsigned char *p) { uint32_t v; std::memcpy(&v, p, 4); return v; };
    auto u16 = [](const unsigned char *p) { uint16_t v; std::memcpy(&v, p, 2); return v; };
    auto f64 = [](const unsigned char *p) { double v; std::memcpy(&v, p, 8); return v; };
    std::vector<Pt> inputCloud;
    double timespan = 0.0;
    stamp_shift = 0.0;
    sampleCloud.clear();
    if (n == 0) {
        timespan_out = 0.0;
        return 0;
    }
    auto blank = []() {
        Pt d;
        std::memset(&d, 0, sizeof(d));
        d.d3 = 1.0f;  // PCL_ADD_POINT4D
        return d;
    };
    if (sensor == SENSOR_VELODYNE) {  // :271-302
        inputCloud.resize(n, blank());
        timespan = f32(rec(n - 1) + L.off_time) - f32(rec(0) + L.off_time);  // :279 (float - float)
        for (int i = 0; i < n; i++) {
            Pt &dst = inputCloud[i];
            dst.x = f32(rec(i) + L.off_x);
            dst.y = f32(rec(i) + L.off_y);
            dst.z = f32(rec(i) + L.off_z);
            dst.intensity = f32(rec(i) + L.off_intensity);
            dst.ny = u16(rec(i) + L.off_ring);
            dst.nz = timespan;
            dst.nx = (f32(rec(i) + L.off_time) + timespan) / timespan;  // :296
        }
        timespan = 0.0;  // :299
    } else if (sensor == SENSOR_LIVOX) {  // :306-323
        inputCloud.resize(n, blank());
        timespan = f32(rec(n - 1) + L.off_time);
        for (int i = 0; i < n; i++) {
            Pt &dst = inputCloud[i];
            dst.x = f32(rec(i) + L.off_x);
            dst.y = f32(rec(i) + L.off_y);
            dst.z = f32(rec(i) + L.off_z);
            dst.intensity = f32(rec(i) + L.off_intensity);
            dst.ny = u16(rec(i) + L.off_ring);
            dst.nz = timespan;
            dst.nx = f32(rec(i) + L.off_time) / timespan;
        }
    } else if (sensor == SENSOR_OUSTER) {  // :324-347
        inputCloud.resize(n, blank());
        timespan = u32(rec(n - 2) + L.off_time);  // :333
        for (int i = 0; i < n; i++) {
            Pt &dst = inputCloud[i];
            dst.x = f32(rec(i) + L.off_x);
            dst.y = f32(rec(i) + L.off_y);
            dst.z = f32(rec(i) + L.off_z);
            dst.intensity = f32(rec(i) + L.off_intensity);
            dst.ny = u16(rec(i) + L.off_ring);
            dst.nz = timespan * 1e-9f;
            dst.nx = u32(rec(i) + L.off_time) / timespan;
        }
        timespan = timespan * 1e-9f;  // :346
    } else {  // SENSOR_ROBOSENSE, :348-383
        const double t0 = f64(rec(0) + L.off_time);
        timespan = f64(rec(n - 1) + L.off_time) - t0;
        for (int i = 0; i < n; i++) {
            const float x = f32(rec(i) + L.off_x), y = f32(rec(i) + L.off_y), z = f32(rec(i) + L.off_z);
            if (!std::isfinite(x) || !std::isfinite(y) || !std::isfinite(z)) continue;
            Pt dst = blank();
            dst.x = x;
            dst.y = y;
            dst.z = z;
            dst.intensity = *(rec(i) + L.off_intensity);  // uint8_t
            dst.ny = u16(rec(i) + L.off_ring);
            dst.nz = timespan;
            dst.nx = (f64(rec(i) + L.off_time) - t0) / timespan;
            inputCloud.push_back(dst);
        }
        stamp_shift = timespan;  // :383
    }
    timespan_out = timespan;  // timeScanEnd = timeScanCur + timespan, :387
    // samplePointCloud, :425-450
    const int cloudSize = (int)inputCloud.size();
    for (int i = 0; i < cloudSize; ++i) {
        if (i % point_filter_num != 0) continue;
        Pt thisPoint = blank();
        thisPoint.x = inputCloud[i].x;
        thisPoint.y = inputCloud[i].y;
        thisPoint.z = inputCloud[i].z;
        thisPoint.intensity = inputCloud[i].intensity;
        thisPoint.nx = inputCloud[i].nx;
        thisPoint.ny = inputCloud[i].ny;
        thisPoint.nz = inputCloud[i].nz;
        float range = pointDistance(thisPoint);
        if (range < lidarMinRange || range > lidarMaxRange) continue;
        sampleCloud.push_back(thisPoint);
    }
    return (int)sampleCloud.size();
}

struct ImuSample {
    double t;
    double acc[3];
    double gyr[3];
};
// eskf_lio/msg/Pose6D.msg
struct Pose6D {
    double offset_time;
    double acc[3], gyr[3], vel[3], pos[3], rot[9];
};
static const double G_m_s2 = 9.8099;  // common_lib.h:23
static const int MAX_INI_COUNT = 100;  // IMU_Processing.hpp:29

struct MeasureGroup {  // common_lib.h:60-71
    double lidar_beg_time = 0.0;
    double observation_end_time = 0.0;
    std::vector<Pt> lidar;
    std::vector<ImuSample> imu;
};

// ImuProcess: IMU_Processing.hpp:34-427
struct ImuProcess {
    bool b_first_frame_ = true;
    bool imu_need_init_ = true;
    int init_iter_num = 1;
    V3 mean_acc = V3(0, 0, -1.0), mean_gyr;
    V3 cov_acc = V3(0.1, 0.1, 0.1), cov_gyr = V3(0.1, 0.1, 0.1);
    V3 angvel_last, acc_s_last;
    M3 Lidar_R_wrt_IMU = M3::I();
    V3 Lidar_T_wrt_IMU;
    ImuSample last_imu_;
    // IMU_Processing.hpp:87 leaves this uninitialised; a fresh heap page reads as 0.
    double last_observation_end_time_ = 0.0;
    std::vector<Pose6D> IMUpose;
    int first_point_repeat = 0;  // diagnostics: how often the begin() point was re-compensated

    ImuProcess() { std::memset(&last_imu_, 0, sizeof(last_imu_)); }

    void Reset() {  // IMU_Processing.hpp:117-141
        angvel_last = V3();
        cov_acc = V3(0.1, 0.1, 0.1);
        cov_gyr = V3(0.1, 0.1, 0.1);
        mean_acc = V3(0, 0, -1.0);
        mean_gyr = V3();
        imu_need_init_ = true;
        b_first_frame_ = true;
        init_iter_num = 1;
        std::memset(&last_imu_, 0, sizeof(last_imu_));
        IMUpose.clear();
    }
    void set_extrinsic(const V3 &t, const M3 &r) {
        Lidar_T_wrt_IMU = t;
        Lidar_R_wrt_IMU = r;
    }

    // IMU_Processing.hpp:161-202
    void IMU_Initial(const MeasureGroup &meas, State &st, int &N) {
        if (b_first_frame_) {
            Reset();
            N = 1;
            b_first_frame_ = false;
            const ImuSample &f = meas.imu.front();
            mean_acc = V3(f.acc[0], f.acc[1], f.acc[2]);
            mean_gyr = V3(f.gyr[0], f.gyr[1], f.gyr[2]);
        }
        for (const ImuSample &imu : meas.imu) {
            V3 cur_acc(imu.acc[0], imu.acc[1], imu.acc[2]), cur_gyr(imu.gyr[0], imu.gyr[1], imu.gyr[2]);
            mean_acc = mean_acc + (cur_acc - mean_acc) / N;
            mean_gyr = mean_gyr + (cur_gyr - mean_gyr) / N;
            for (int a = 0; a < 3; a++) {
                double da = cur_acc[a] - mean_acc[a], dg = cur_gyr[a] - mean_gyr[a];
                cov_acc[a] = cov_acc[a] * (N - 1.0) / N + da * da * (N - 1.0) / (N * N);
                cov_gyr[a] = cov_gyr[a] * (N - 1.0) / N + dg * dg * (N - 1.0) / (N * N);
            }
            N++;
        }
        st.gravity = (mean_acc * -1.0) / norm(mean_acc) * G_m_s2;
        st.bias_g = mean_gyr;
        st.R_L_I = Lidar_R_wrt_IMU;
        st.T_L_I = Lidar_T_wrt_IMU;
        last_imu_ = meas.imu.back();
    }

    // Forward propagation part of UndistortPcl: IMU_Processing.hpp:204-330.
    // Fills IMUpose and updates state_inout (cov always; end pose unless EKF_stop_flg).
    void Propagate(const MeasureGroup &meas, State &st, bool EKF_stop_flg) {
        std::vector<ImuSample> v_imu;
        v_imu.push_back(last_imu_);
        v_imu.insert(v_imu.end(), meas.imu.begin(), meas.imu.end());
        const double imu_end_time = v_imu.back().t;
        const double pcl_beg_time = meas.lidar_beg_time;
        const double pcl_end_time = meas.observation_end_time;

        IMUpose.clear();
        Pose6D p0;
        p0.offset_time = 0.0;
        for (int i = 0; i < 3; i++) {
            p0.acc[i] = acc_s_last[i];
            p0.gyr[i] = angvel_last[i];
            p0.vel[i] = st.vel_end[i];
            p0.pos[i] = st.pos_end[i];
        }
        std::memcpy(p0.rot, st.rot_end.m, sizeof(p0.rot));
        IMUpose.push_back(p0);

        V3 acc_imu, angvel_avr, acc_avr, vel_imu(st.vel_end), pos_imu(st.pos_end);
        M3 R_imu(st.rot_end);
        double dt = 0;
        for (size_t k = 0; k + 1 < v_imu.size(); k++) {
            const ImuSample &head = v_imu[k];
            const ImuSample &tail = v_imu[k + 1];
            if (tail.t < last_observation_end_time_) continue;
            for (int a = 0; a < 3; a++) {
                angvel_avr[a] = 0.5 * (head.gyr[a] + tail.gyr[a]);
                acc_avr[a] = 0.5 * (head.acc[a] + tail.acc[a]);
            }
            angvel_avr = angvel_avr - st.bias_g;
            acc_avr = acc_avr * G_m_s2 / norm(mean_acc) - st.bias_a;
            if (head.t < last_observation_end_time_)
                dt = tail.t - last_observation_end_time_;
            else
                dt = tail.t - head.t;

            /* covariance propagation, IMU_Processing.hpp:262-288 */
            M3 acc_avr_skew = skew(acc_avr);
            M3 Exp_f = Exp(angvel_avr, dt);
            Mat F((size_t)DIM * DIM, 0.0), Q((size_t)DIM * DIM, 0.0);
            for (int i = 0; i < DIM; i++) F[(size_t)i * DIM + i] = 1.0;
            auto setblk = [&](Mat &M, int r0, int c0, const M3 &B) {
                for (int i = 0; i < 3; i++)
                    for (int j = 0; j < 3; j++) M[(size_t)(r0 + i) * DIM + c0 + j] = B(i, j);
            };
            setblk(F, 0, 0, Exp(angvel_avr, -dt));
            setblk(F, 0, 15, M3::I() * (-dt));
            setblk(F, 3, 12, M3::I() * dt);
            setblk(F, 12, 0, (R_imu * -1.0) * acc_avr_skew * dt);
            setblk(F, 12, 18, R_imu * (-dt));
            setblk(F, 12, 21, M3::I() * dt);
            M3 cad, cgd;
            for (int a = 0; a < 3; a++) {
                cad(a, a) = cov_acc[a];
                cgd(a, a) = cov_gyr[a];
            }
            for (int a = 0; a < 3; a++) Q[(size_t)a * DIM + a] = cov_gyr[a] * dt * dt * 10000;
            setblk(Q, 3, 3, R_imu * cgd * transpose(R_imu) * dt * dt * 10000);
            setblk(Q, 12, 12, R_imu * cad * transpose(R_imu) * dt * dt * 10000);
            for (int a = 0; a < 3; a++) {
                Q[(size_t)(15 + a) * DIM + 15 + a] = 0.0001 * dt * dt;
                Q[(size_t)(18 + a) * DIM + 18 + a] = 0.0001 * dt * dt;
            }
            Mat P(st.cov, st.cov + DIM * DIM);
            Mat FP = mat_mul(F, DIM, DIM, P, DIM);
            Mat FPFt = mat_mul(FP, DIM, DIM, mat_T(F, DIM, DIM), DIM);
            for (int i = 0; i < DIM * DIM; i++) st.cov[i] = FPFt[i] + Q[i];

            R_imu = R_imu * Exp_f;
            acc_imu = R_imu * acc_avr + st.gravity;
            pos_imu = pos_imu + vel_imu * dt + acc_imu * 0.5 * dt * dt;
            vel_imu = vel_imu + acc_imu * dt;
            angvel_last = angvel_avr;
            acc_s_last = acc_imu;
            Pose6D p;
            p.offset_time = tail.t - pcl_beg_time;
            for (int i = 0; i < 3; i++) {
                p.acc[i] = acc_imu[i];
                p.gyr[i] = angvel_avr[i];
                p.vel[i] = vel_imu[i];
                p.pos[i] = pos_imu[i];
            }
            std::memcpy(p.rot, R_imu.m, sizeof(p.rot));
            IMUpose.push_back(p);
        }
        dt = pcl_end_time - imu_end_time;
        if (!EKF_stop_flg) {  // IMU_Processing.hpp:314-319
            st.vel_end = vel_imu + acc_imu * dt;
            st.rot_end = R_imu * Exp(angvel_avr, dt);
            st.pos_end = pos_imu + vel_imu * dt + acc_imu * 0.5 * dt * dt;
        }
        last_observation_end_time_ = pcl_end_time;
    }

    // Backward per-point compensation: IMU_Processing.hpp:332-370 (pcl sorted by normal_x).
    void Compensate(const State &st, std::vector<Pt> &pcl) {
        first_point_repeat = 0;
        if (pcl.empty() || IMUpose.size() < 2) return;
        M3 RLI_T = transpose(st.R_L_I), Rend_T = transpose(st.rot_end);
        long it = (long)pcl.size() - 1;
        for (long kp = (long)IMUpose.size() - 1; kp != 0; kp--) {
            const Pose6D &head = IMUpose[kp - 1];
            M3 R_imu;
            std::memcpy(R_imu.m, head.rot, sizeof(head.rot));
            V3 acc_imu(head.acc[0], head.acc[1], head.acc[2]);
            V3 vel_imu(head.vel[0], head.vel[1], head.vel[2]);
            V3 pos_imu(head.pos[0], head.pos[1], head.pos[2]);
            V3 angvel(head.gyr[0], head.gyr[1], head.gyr[2]);
            for (; (double)(pcl[it].nx * pcl[it].nz) > head.offset_time; it--) {
                double dt = (double)(pcl[it].nx * pcl[it].nz) - head.offset_time;
                M3 R_i = R_imu * Exp(angvel, dt);
                V3 T_ei = pos_imu + vel_imu * dt + acc_imu * 0.5 * dt * dt - st.pos_end;
                V3 P_i(pcl[it].x, pcl[it].y, pcl[it].z);
                V3 Pc = RLI_T * (Rend_T * (R_i * (st.R_L_I * P_i + st.T_L_I) + T_ei) - st.T_L_I);
                pcl[it].x = (float)Pc[0];
                pcl[it].y = (float)Pc[1];
                pcl[it].z = (float)Pc[2];
                if (it == 0) {
                    first_point_repeat++;
                    break;
                }
            }
        }
    }

    // UndistortPcl: IMU_Processing.hpp:204-371
    void UndistortPcl(const MeasureGroup &meas, State &st, std::vector<Pt> &pcl_out, bool EKF_stop_flg) {
        pcl_out = meas.lidar;
        std::sort(pcl_out.begin(), pcl_out.end(), [](const Pt &x, const Pt &y) { return x.nx < y.nx; });
        Propagate(meas, st, EKF_stop_flg);
        Compensate(st, pcl_out);
    }

    // Process: IMU_Processing.hpp:373-427.  Returns true when pcl_out was produced.
    bool Process(const MeasureGroup &meas, State &st, std::vector<Pt> &pcl_out, bool EKF_stop_flg) {
        if (meas.imu.empty()) return false;
        if (imu_need_init_) {
            IMU_Initial(meas, st, init_iter_num);
            imu_need_init_ = true;
            last_imu_ = meas.imu.back();
            if (init_iter_num > MAX_INI_COUNT) {
                imu_need_init_ = false;
                cov_acc = V3(0.1, 0.1, 0.1);
                cov_gyr = V3(0.1, 0.1, 0.1);
            }
            return false;
        }
        st.gravity = V3(0, 0, -9.801);  // IMU_Processing.hpp:408-412
        st.bias_a = V3(1.44e-04, 1.44e-04, 1.44e-04);
        st.bias_g = V3(5.3e-05, 5.3e-05, 5.3e-05);
        st.R_L_I = Lidar_R_wrt_IMU;
        st.T_L_I = Lidar_T_wrt_IMU;
        UndistortPcl(meas, st, pcl_out, EKF_stop_flg);
        last_imu_ = meas.imu.back();
        return true;
    }
};

// ======================================================================== map back-ends
struct Box {
    float mn[3], mx[3];
};

struct MapBackend {
    virtual ~MapBackend() {}
    virtual void build(const std::vector<P4> &pts) = 0;
    virtual bool has_root() const = 0;
    // k nearest, ascending; d2 float as ikd_Tree.cpp:1682
    virtual void knn(float qx, float qy, float qz, int k, std::vector<P4> &near, std::vector<float> &d2) = 0;
    virtual int add_points(const std::vector<P4> &pts, bool downsample_on) = 0;
    virtual int delete_boxes(const std::vector<Box> &boxes) = 0;
    virtual int validnum() = 0;
    virtual void flatten(std::vector<P4> &out) = 0;
};

// ---- PortMap: independent restatement of the ikd-Tree's contents + query semantics
//   contents:  Build (no dedupe)                      ikd_Tree.cpp:408-423
//              Add_Points downsample-on-insert         ikd_Tree.cpp:477-522 (sequential)
//              Add_Points raw                          ikd_Tree.cpp:549-554
//              Delete_Point_Boxes, [min,max) boxes     ikd_Tree.cpp:631-658, 796
//   query:     exact k nearest by calc_dist, ascending ikd_Tree.cpp:425-461, 1061-1244
//   tie-break: (d2 float, x, y, z) -- the reference's order under exact d2 ties depends on
//              tree shape (strict '<' at ikd_Tree.cpp:1088,1099); parity is asserted only
//              where d2[k-1] != d2[k].
struct PortMap : MapBackend {
    float ds;
    float cell;  // search cell edge
    struct Rec {
        P4 p;
        bool alive;
    };
    std::vector<Rec> pts;
    std::unordered_map<uint64_t, std::vector<int>> grid;
    int live = 0;
    bool built = false;

    explicit PortMap(float downsample_size) : ds(downsample_size), cell(2.0f * downsample_size) {}

    static uint64_t key(int64_t cx, int64_t cy, int64_t cz) {
        return ((uint64_t)(cx & 0x1FFFFF) << 42) | ((uint64_t)(cy & 0x1FFFFF) << 21) | (uint64_t)(cz & 0x1FFFFF);
    }
    int64_t cidx(float v) const { return (int64_t)std::floor((double)v / (double)cell); }
    void insert_raw(const P4 &p) {
        int id = (int)pts.size();
        pts.push_back({p, true});
        grid[key(cidx(p.x), cidx(p.y), cidx(p.z))].push_back(id);
        live++;
    }
    void build(const std::vector<P4> &in) override {
        pts.clear();
        grid.clear();
        live = 0;
        for (const P4 &p : in) insert_raw(p);
        built = !in.empty();
    }
    bool has_root() const override { return built; }
    int validnum() override { return live; }
    void flatten(std::vector<P4> &out) override {
        out.clear();
        for (const Rec &r : pts)
            if (r.alive) out.push_back(r.p);
    }
    static bool cand_less(float da, const P4 &a, float db, const P4 &b) {
        if (da != db) return da < db;
        if (a.x != b.x) return a.x < b.x;
        if (a.y != b.y) return a.y < b.y;
        return a.z < b.z;
    }
    void knn(float qx, float qy, float qz, int k, std::vector<P4> &near, std::vector<float> &d2) override {
        near.clear();
        d2.clear();
        struct C {
            float d;
            int id;
        };
        std::vector<C> best;  // sorted ascending, size <= k
        auto consider = [&](int id) {
            const P4 &p = pts[id].p;
            float d = calc_dist(qx, qy, qz, p.x, p.y, p.z);
            if ((int)best.size() == k && !cand_less(d, p, best.back().d, pts[best.back().id].p)) return;
            C c{d, id};
            auto it = best.begin();
            while (it != best.end() && !cand_less(d, p, it->d, pts[it->id].p)) ++it;
            best.insert(it, c);
            if ((int)best.size() > k) best.pop_back();
        };
        int64_t cx = cidx(qx), cy = cidx(qy), cz = cidx(qz);
        const int RMAX = 6;
        bool done = false;
        for (int R = 0; R <= RMAX && !done; R++) {
            for (int64_t ix = cx - R; ix <= cx + R; ix++)
                for (int64_t iy = cy - R; iy <= cy + R; iy++)
                    for (int64_t iz = cz - R; iz <= cz + R; iz++) {
                        if (std::max(std::max(std::llabs(ix - cx), std::llabs(iy - cy)), std::llabs(iz - cz)) != R) continue;
                        auto it = grid.find(key(ix, iy, iz));
                        if (it == grid.end()) continue;
                        for (int id : it->second)
                            if (pts[id].alive) consider(id);
                    }
            if ((int)best.size() == k) {
                // every unseen point is farther than R*cell (minus rounding slack) on some axis
                double cov = (double)R * (double)cell;
                double slack = 1e-4 * (1.0 + std::fabs(qx) + std::fabs(qy) + std::fabs(qz)) * 1e-2;
                double lim = cov - slack;
                if (lim > 0 && (double)best.back().d < lim * lim * (1.0 - 1e-5)) done = true;
            }
        }
        if (!done) {  // exact fallback
            best.clear();
            for (int id = 0; id < (int)pts.size(); id++)
                if (pts[id].alive) consider(id);
        }
        for (const C &c : best) {
            near.push_back(pts[c.id].p);
            d2.push_back(c.d);
        }
    }
    void box_ids(const Box &b, std::vector<int> &ids) {
        ids.clear();
        int64_t lo[3], hi[3];
        for (int a = 0; a < 3; a++) {
            lo[a] = cidx(b.mn[a]) - 1;
            hi[a] = cidx(b.mx[a]) + 1;
        }
        double vol = (double)(hi[0] - lo[0] + 1) * (double)(hi[1] - lo[1] + 1) * (double)(hi[2] - lo[2] + 1);
        auto inside = [&](const P4 &p) {
            return b.mn[0] <= p.x && b.mx[0] > p.x && b.mn[1] <= p.y && b.mx[1] > p.y && b.mn[2] <= p.z && b.mx[2] > p.z;
        };
        if (vol > 4096.0) {
            for (int id = 0; id < (int)pts.size(); id++)
                if (pts[id].alive && inside(pts[id].p)) ids.push_back(id);
            return;
        }
        for (int64_t ix = lo[0]; ix <= hi[0]; ix++)
            for (int64_t iy = lo[1]; iy <= hi[1]; iy++)
                for (int64_t iz = lo[2]; iz <= hi[2]; iz++) {
                    auto it = grid.find(key(ix, iy, iz));
                    if (it == grid.end()) continue;
                    for (int id : it->second)
                        if (pts[id].alive && inside(pts[id].p)) ids.push_back(id);
                }
        std::sort(ids.begin(), ids.end());
    }
    void kill(int id) {
        if (pts[id].alive) {
            pts[id].alive = false;
            live--;
        }
    }
    int add_points(const std::vector<P4> &in, bool downsample_on) override {
        int counter = 0;
        std::vector<int> ids;
        for (const P4 &p : in) {
            if (!downsample_on) {
                insert_raw(p);
                continue;
            }
            Box b;  // ikd_Tree.cpp:491-499
            const float pv[3] = {p.x, p.y, p.z};
            float mid[3];
            for (int a = 0; a < 3; a++) {
                b.mn[a] = std::floor(pv[a] / ds) * ds;
                b.mx[a] = b.mn[a] + ds;
                mid[a] = (float)((double)b.mn[a] + (double)(b.mx[a] - b.mn[a]) / 2.0);
            }
            box_ids(b, ids);
            float min_dist = calc_dist(p.x, p.y, p.z, mid[0], mid[1], mid[2]);
            int win = -1;  // -1: the new point
            for (int id : ids) {
                const P4 &e = pts[id].p;
                float d = calc_dist(e.x, e.y, e.z, mid[0], mid[1], mid[2]);
                bool better = d < min_dist;
                // among existing points at exactly the same distance the reference keeps the first
                // in tree-traversal order; here: smallest (x,y,z)
                if (!better && win >= 0 && d == min_dist && cand_less(d, e, d, pts[win].p)) better = true;
                if (better) {
                    min_dist = d;
                    win = id;
                }
            }
            P4 res = (win < 0) ? p : pts[win].p;
            bool same = std::fabs(p.x - res.x) < 1e-6 && std::fabs(p.y - res.y) < 1e-6 && std::fabs(p.z - res.z) < 1e-6;
            if (ids.size() > 1 || same) {  // ikd_Tree.cpp:515-521
                for (int id : ids) kill(id);
                insert_raw(res);
                counter++;
            }
        }
        built = built || live > 0;
        return counter;
    }
    int delete_boxes(const std::vector<Box> &boxes) override {
        int c = 0;
        std::vector<int> ids;
        for (const Box &b : boxes) {
            box_ids(b, ids);
            for (int id : ids) {
                kill(id);
                c++;
            }
        }
        return c;
    }
};

// ---- RefMap: the unmodified reference ikd-Tree through oracle/_ref/libikd_ref.so
// (function pointers are resolved by the C API layer with dlopen so that liboracle.so
//  itself has no link-time dependency on the reference build).
struct RefApi {
    void *(*create)(float) = nullptr;
    void (*destroy)(void *) = nullptr;
    void (*build)(void *, const float *, int) = nullptr;
    void (*knn)(void *, const float *, int, int, float *, float *, int *) = nullptr;
    int (*add)(void *, const float *, int, int) = nullptr;
    int (*delete_boxes)(void *, const float *, int) = nullptr;
    int (*validnum)(void *) = nullptr;
    int (*size)(void *) = nullptr;
    int (*has_root)(void *) = nullptr;
    int (*flatten)(void *, float *, int) = nullptr;
    bool ok() const { return create && knn && add; }
};
struct RefMap : MapBackend {
    const RefApi *api;
    void *h;
    RefMap(const RefApi *a, float ds) : api(a) { h = api->create(ds); }
    ~RefMap() override { api->destroy(h); }
    void build(const std::vector<P4> &pts) override { api->build(h, &pts[0].x, (int)pts.size()); }
    bool has_root() const override { return api->has_root(h) != 0; }
    void knn(float qx, float qy, float qz, int k, std::vector<P4> &near, std::vector<float> &d2) override {
        float q[3] = {qx, qy, qz};
        std::vector<float> op((size_t)k * 4), od(k);
        int cnt = 0;
        api->knn(h, q, 1, k, op.data(), od.data(), &cnt);
        near.resize(cnt);
        d2.resize(cnt);
        for (int j = 0; j < cnt; j++) {
            near[j] = {op[4 * j], op[4 * j + 1], op[4 * j + 2], op[4 * j + 3]};
            d2[j] = od[j];
        }
    }
    int add_points(const std::vector<P4> &pts, bool ds_on) override {
        if (pts.empty()) return 0;
        return api->add(h, &pts[0].x, (int)pts.size(), ds_on ? 1 : 0);
    }
    int delete_boxes(const std::vector<Box> &boxes) override {
        if (boxes.empty()) return 0;
        return api->delete_boxes(h, &boxes[0].mn[0], (int)boxes.size());
    }
    int validnum() override { return api->validnum(h); }
    void flatten(std::vector<P4> &out) override {
        int n = api->size(h) + 16;
        std::vector<float> buf((size_t)n * 4);
        int m = api->flatten(h, buf.data(), n);
        if (m > n) {
            buf.resize((size_t)m * 4);
            m = api->flatten(h, buf.data(), m);
        }
        out.resize(m);
        for (int i = 0; i < m; i++) out[i] = {buf[4 * i], buf[4 * i + 1], buf[4 * i + 2], buf[4 * i + 3]};
    }
};

// ======================================================================== the per-scan update
struct LioConfig {
    int max_iteration = 4;           // mapping/max_iteration (feat.yaml:46: 10; BASELINE configs: 4)
    double filter_size_surf = 0.5;   // feat.yaml:47
    double filter_size_map = 0.5;    // feat.yaml:48
    double cube_len = 1000.0;        // feat.yaml:49
    bool extrinsic_est_en = false;   // feat.yaml:50
    int featptsThreshold = 30;       // feat.yaml:7
    double beta = 0.1;               // feat.yaml:8
    float det_range = 300.0f;        // laserMapping.cpp:304
    V3 extrinT;                      // feat.yaml:51
    M3 extrinR = M3::I();            // feat.yaml:52
};

struct ThermalInputs {  // what the tis / edge callbacks leave behind (laserMapping.cpp:471-498)
    bool tis_online = false;
    int recv_n = 0;
    // g_tis_odom_delta (pos, quat wxyz, vel) and ..._lframe2lframe
    double delta_pos[3] = {0, 0, 0}, delta_quat[4] = {1, 0, 0, 0}, delta_vel[3] = {0, 0, 0};
    double cov_slots[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double l2l_pos[3] = {0, 0, 0}, l2l_quat[4] = {1, 0, 0, 0}, l2l_vel[3] = {0, 0, 0};
    double l2l_cov_slots[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

struct IterRecord {  // one row of Log/mat_out.txt (laserMapping.cpp:936-937) plus the algebra
    int iter = 0;
    int effct_feat_num = 0;
    double total_residual = 0.0;
    double res_mean_last = 0.0;
    bool converged = false;
    bool ekf_stop = false;
    bool did_match = false;
    int n_down = 0;
    double HtH[144];       // 12x12 row-major
    double Htr[12];        // H^T * meas_vec
    double pose_in[24];    // rot_end, pos_end, R_L_I, T_L_I the residuals were evaluated at
    double state_out[36];  // state after this iteration's update
    double solution[24];
};

struct ScanResult {
    bool had_points = false;      // feats_undistort non-empty
    bool built_map = false;       // this scan initialised the map
    bool did_update = false;      // featsFromMapNum >= 5 and the loop ran
    int n_raw = 0, n_down = 0, map_points_before = 0;
    int deleted = 0, added = 0;
    bool ekf_stop = false;
    std::vector<IterRecord> iters;
    std::vector<Pt> feats_undistort, feats_down;
    std::vector<P4> nearest;       // n_down x 5 (last match pass), padded with zeros
    std::vector<int> nearest_cnt;  // n_down
    std::vector<float> nearest_d2;  // n_down x 5
    std::vector<uint8_t> selected; // n_down, point_selected_surf after the last iteration
    std::vector<P4> added_ds, added_raw;
    double eigvals[6], eigvecs[36];
    double state_prop[36];
};

inline State odomToState(const double pos[3], const double q[4], const double vel[3], const double cs[8]) {
    // odomToStateGruop, laserMapping.cpp:218-239
    State r;
    r.pos_end = V3(pos[0], pos[1], pos[2]);
    r.rot_end = quat_to_rot(q[0], q[1], q[2], q[3]);
    r.vel_end = V3(vel[0], vel[1], vel[2]);
    r.bias_g = V3(cs[4], cs[5], cs[6]);
    r.bias_a = V3(cs[1], cs[2], cs[3]);
    r.gravity = V3(0, 0, cs[7]);
    return r;
}

struct Lio {
    LioConfig cfg;
    std::unique_ptr<MapBackend> map;
    ImuProcess imu;
    State state, last_nodegared_state, last_state;
    bool flg_first_scan = true;
    double first_lidar_time = 0.0;
    bool EKF_stop_flg = false;
    bool flg_EKF_inited = false;
    std::deque<int> effct_q;                          // effct_feat_numQueue, laserMapping.cpp:192-193
    int dynamic_effect_featurepoints_threshold = 100;  // laserMapping.cpp:97
    long lidar_frame_counter_num = 0;
    long lidar_cnt = 0;
    double zeta_l = 0.0, zeta_t = 0.0;
    bool Localmap_Initialized = false;
    Box LocalMap_Points;
    std::vector<std::vector<P4>> Nearest_Points;
    int omp_threads = 1;  // 1 = as shipped (pragmas commented out, laserMapping.cpp:827-828)
    // per-stage wall clock of the last scan (seconds)
    double t_deskew = 0, t_voxel = 0, t_knn = 0, t_resid = 0, t_solve = 0, t_insert = 0, t_delete = 0;

    Lio(const LioConfig &c, MapBackend *m) : cfg(c), map(m) { imu.set_extrinsic(cfg.extrinT, cfg.extrinR); }

    // feat_points_cbk bookkeeping: laserMapping.cpp:424-446
    void on_lidar_msg() {
        lidar_frame_counter_num++;
        if (lidar_frame_counter_num > 100) dynamic_effect_featurepoints_threshold = cfg.featptsThreshold;
        lidar_cnt++;
    }
    // tn_cbk: laserMapping.cpp:491-498
    void on_edge_count(int recv_n) {
        double alpha_t = recv_n / 307200.0;
        zeta_t = 2.0 / (1.0 + std::exp(-alpha_t)) - 1;
    }

    static void body_to_world(const State &s, float x, float y, float z, float out[3]) {
        // pointBodyToWorld, laserMapping.cpp:260-269 / :835-840
        V3 pb(x, y, z);
        V3 pg = s.rot_end * (s.R_L_I * pb + s.T_L_I) + s.pos_end;
        out[0] = (float)pg[0];
        out[1] = (float)pg[1];
        out[2] = (float)pg[2];
    }

    // lasermap_fov_segment: laserMapping.cpp:313-369
    int fov_segment(const V3 &pos_LiD, std::vector<Box> &cub_needrm) {
        cub_needrm.clear();
        const float MOV_THRESHOLD = 1.5f;
        const float DET_RANGE = cfg.det_range;
        if (!Localmap_Initialized) {
            for (int i = 0; i < 3; i++) {
                LocalMap_Points.mn[i] = (float)(pos_LiD[i] - cfg.cube_len / 2.0);
                LocalMap_Points.mx[i] = (float)(pos_LiD[i] + cfg.cube_len / 2.0);
            }
            Localmap_Initialized = true;
            return 0;
        }
        float dist_to_map_edge[3][2];
        bool need_move = false;
        for (int i = 0; i < 3; i++) {
            dist_to_map_edge[i][0] = (float)std::fabs(pos_LiD[i] - (double)LocalMap_Points.mn[i]);
            dist_to_map_edge[i][1] = (float)std::fabs(pos_LiD[i] - (double)LocalMap_Points.mx[i]);
            if (dist_to_map_edge[i][0] <= MOV_THRESHOLD * DET_RANGE || dist_to_map_edge[i][1] <= MOV_THRESHOLD * DET_RANGE)
                need_move = true;
        }
        if (!need_move) return 0;
        Box New = LocalMap_Points, tmp;
        float mov_dist = (float)std::max((cfg.cube_len - 2.0 * MOV_THRESHOLD * DET_RANGE) * 0.5 * 0.9,
                                         double(DET_RANGE * (MOV_THRESHOLD - 1)));
        for (int i = 0; i < 3; i++) {
            tmp = LocalMap_Points;
            if (dist_to_map_edge[i][0] <= MOV_THRESHOLD * DET_RANGE) {
                New.mx[i] -= mov_dist;
                New.mn[i] -= mov_dist;
                tmp.mn[i] = LocalMap_Points.mx[i] - mov_dist;
                cub_needrm.push_back(tmp);
            } else if (dist_to_map_edge[i][1] <= MOV_THRESHOLD * DET_RANGE) {
                New.mx[i] += mov_dist;
                New.mn[i] += mov_dist;
                tmp.mx[i] = LocalMap_Points.mn[i] + mov_dist;
                cub_needrm.push_back(tmp);
            }
        }
        LocalMap_Points = New;
        if (!cub_needrm.empty()) return map->delete_boxes(cub_needrm);
        return 0;
    }

    // map_incremental: laserMapping.cpp:582-630
    void map_incremental(const std::vector<Pt> &feats_down, ScanResult &res) {
        const int NUM_MATCH_POINTS = 5;
        const double fs = cfg.filter_size_map;
        std::vector<P4> PointToAdd, PointNoNeedDownsample;
        int n = (int)feats_down.size();
        for (int i = 0; i < n; i++) {
            float w[3];
            body_to_world(state, feats_down[i].x, feats_down[i].y, feats_down[i].z, w);
            P4 pw{w[0], w[1], w[2], feats_down[i].intensity};
            if (!Nearest_Points[i].empty() && flg_EKF_inited) {
                const std::vector<P4> &points_near = Nearest_Points[i];
                bool need_add = true;
                float mid[3];
                for (int a = 0; a < 3; a++) mid[a] = (float)(std::floor((double)w[a] / fs) * fs + 0.5 * fs);
                float dist = calc_dist(w[0], w[1], w[2], mid[0], mid[1], mid[2]);
                if ((double)std::fabs(points_near[0].x - mid[0]) > 0.5 * fs && (double)std::fabs(points_near[0].y - mid[1]) > 0.5 * fs &&
                    (double)std::fabs(points_near[0].z - mid[2]) > 0.5 * fs) {
                    PointNoNeedDownsample.push_back(pw);
                    continue;
                }
                for (int j = 0; j < NUM_MATCH_POINTS; j++) {
                    if ((int)points_near.size() < NUM_MATCH_POINTS) break;
                    if (calc_dist(points_near[j].x, points_near[j].y, points_near[j].z, mid[0], mid[1], mid[2]) < dist) {
                        need_add = false;
                        break;
                    }
                }
                if (need_add) PointToAdd.push_back(pw);
            } else {
                PointToAdd.push_back(pw);
            }
        }
        map->add_points(PointToAdd, true);
        map->add_points(PointNoNeedDownsample, false);
        res.added = (int)(PointToAdd.size() + PointNoNeedDownsample.size());
        res.added_ds = PointToAdd;
        res.added_raw = PointNoNeedDownsample;
    }

    // One pass of `while (sync_packages(Measures))` body: laserMapping.cpp:731-1177
    // (publishing and the O(map) flatten at :1170-1175 are outside the hot path).
    void process_scan(const MeasureGroup &meas, const ThermalInputs &th, ScanResult &res);
};

double now_sec();

}  // namespace orc
