// daliti_b200/csrc/host/eskf_lio_host.hpp
//
// Host side of the drop-in: the parts of eskf_lio's per-scan update that are sequential and
// tiny stay in C++ on the CPU, mirroring the reference's types and control flow
//   StatesGroup and its (+) / (-) / scale algebra        eskf_lio/include/common_lib.h:73-228
//   SO(3) Exp / Log                                      eskf_lio/include/so3_math.h:11-81
//   ImuProcess (initialisation + forward propagation)    eskf_lio/src/IMU_Processing.hpp:97-330, 373-427
//   the iteration control, Kalman algebra, degradation
//   window, zeta blend, local-map cube                   eskf_lio/src/laserMapping.cpp:313-369, 731-1177
// while everything per point goes through the device C ABI (include/daliti_b200.h).
//
// The Kalman algebra is written in the reduced form the device outputs allow
// (SURVEY.md section 8b): with K_1 = (H^T H + (P/R)^-1)^-1,
//     solution = K_1[:, :12] (H^T r - H^T H vec_12) + vec,     G = K_1[:, :12] H^T H.
#pragma once
#include <cmath>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

#include "../../../include/daliti_b200_lio.h"

namespace dlt_host {

constexpr int kDim = 24;  // DIM_OF_STATES, common_lib.h:24

struct Vec3 {
    double x = 0, y = 0, z = 0;
    Vec3() {}
    Vec3(double a, double b, double c) : x(a), y(b), z(c) {}
    explicit Vec3(const double *p) : x(p[0]), y(p[1]), z(p[2]) {}
    Vec3 operator+(const Vec3 &o) const { return Vec3(x + o.x, y + o.y, z + o.z); }
    Vec3 operator-(const Vec3 &o) const { return Vec3(x - o.x, y - o.y, z - o.z); }
    Vec3 operator*(double s) const { return Vec3(x * s, y * s, z * s); }
    Vec3 operator/(double s) const { return Vec3(x / s, y / s, z / s); }
    double norm() const { return std::sqrt(x * x + y * y + z * z); }
    void store(double *p) const {
        p[0] = x;
        p[1] = y;
        p[2] = z;
    }
};

struct Mat3 {  // row-major
    double a[9];
    Mat3() { std::memset(a, 0, sizeof(a)); }
    static Mat3 identity() {
        Mat3 m;
        m.a[0] = m.a[4] = m.a[8] = 1.0;
        return m;
    }
    static Mat3 from(const double *p) {
        Mat3 m;
        std::memcpy(m.a, p, sizeof(m.a));
        return m;
    }
    static Mat3 hat(const Vec3 &v) {  // SKEW_SYM_MATRX, so3_math.h:9
        Mat3 m;
        m.a[1] = -v.z;
        m.a[2] = v.y;
        m.a[3] = v.z;
        m.a[5] = -v.x;
        m.a[6] = -v.y;
        m.a[7] = v.x;
        return m;
    }
    Mat3 operator*(const Mat3 &o) const {
        Mat3 r;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) r.a[3 * i + j] = a[3 * i] * o.a[j] + a[3 * i + 1] * o.a[3 + j] + a[3 * i + 2] * o.a[6 + j];
        return r;
    }
    Vec3 operator*(const Vec3 &v) const {
        return Vec3(a[0] * v.x + a[1] * v.y + a[2] * v.z, a[3] * v.x + a[4] * v.y + a[5] * v.z, a[6] * v.x + a[7] * v.y + a[8] * v.z);
    }
    Mat3 operator*(double s) const {
        Mat3 r;
        for (int i = 0; i < 9; i++) r.a[i] = a[i] * s;
        return r;
    }
    Mat3 operator+(const Mat3 &o) const {
        Mat3 r;
        for (int i = 0; i < 9; i++) r.a[i] = a[i] + o.a[i];
        return r;
    }
    Mat3 t() const {
        Mat3 r;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) r.a[3 * i + j] = a[3 * j + i];
        return r;
    }
    double trace() const { return a[0] + a[4] + a[8]; }
};

Mat3 so3_exp_rate(const Vec3 &ang_vel, double dt);       // Exp(ang_vel, dt), so3_math.h:32-52
Mat3 so3_exp(double v1, double v2, double v3);           // Exp(v1, v2, v3), so3_math.h:54-72
Vec3 so3_log(const Mat3 &R);                             // Log(R), so3_math.h:75-81
Mat3 quat_to_rot(double w, double x, double y, double z);

struct StatesGroup {  // common_lib.h:73-228
    Mat3 rot_end = Mat3::identity();
    Vec3 pos_end;
    Mat3 R_L_I = Mat3::identity();
    Vec3 T_L_I;
    Vec3 vel_end, bias_g, bias_a, gravity;
    double cov[kDim * kDim];
    StatesGroup();
    StatesGroup boxplus(const double *add) const;           // operator+(vector)        :115-129
    StatesGroup compose(const StatesGroup &b) const;        // operator+(StatesGroup)   :131-144
    void boxplus_inplace(const double *add);                // operator+=(vector)       :147-158
    void boxminus(const StatesGroup &b, double *out) const; // operator-(StatesGroup)   :173-187
    StatesGroup scaled(double s) const;                     // operator*(double)        :190-205
    void to_flat(double *f612) const;
    void from_flat(const double *f612);
    void pose24(double *p) const;
};

struct ImuSample {
    double t, acc[3], gyr[3];
};
struct Pose6D {  // eskf_lio/msg/Pose6D.msg; same 22-double layout the device expects
    double offset_time, acc[3], gyr[3], vel[3], pos[3], rot[9];
};
static_assert(sizeof(Pose6D) == 22 * sizeof(double), "Pose6D layout");

// ImuProcess without the per-point loop (that part is dlt_scan_deskew)
class ImuProcess {
   public:
    ImuProcess();
    void Reset();                                           // IMU_Processing.hpp:117-141
    void set_extrinsic(const Vec3 &t, const Mat3 &r);       // :155-159
    // Process, :373-427.  Returns true when the scan is to be undistorted (IMUpose / state updated).
    // defer_cov: only the poses are integrated now (that is all the deskew kernel waits for); the covariance propagation of
    // the same IMU steps (:262-288) is recorded and runs in FinishCovariance, which the caller invokes once the scan's first
    // kernels are in flight and before anything reads state.cov.  Same operations on the same numbers either way.
    bool Process(const std::vector<ImuSample> &imu, double lidar_beg_time, double observation_end_time, StatesGroup &state, bool EKF_stop_flg,
                 bool defer_cov = false);
    void FinishCovariance(StatesGroup &state);
    void force_ready(const Vec3 &mean_acc, const ImuSample &last);
    bool need_init() const { return imu_need_init_; }
    std::vector<Pose6D> IMUpose;

   private:
    void IMU_Initial(const std::vector<ImuSample> &imu, StatesGroup &state, int &N);  // :161-202
    void Propagate(const std::vector<ImuSample> &imu, double pcl_beg_time, double pcl_end_time, StatesGroup &state,
                   bool EKF_stop_flg, bool defer_cov);                                  // :204-330
    struct CovStep {  // what one IMU step's F and Q are made of (:270-286)
        Vec3 angvel_avr, acc_avr;
        Mat3 R_imu;
        double dt;
    };
    void PropagateCovStep(const CovStep &s, StatesGroup &state, double *__restrict__ F, double *__restrict__ Q, double *__restrict__ FP) const;
    std::vector<CovStep> pending_cov_;
    bool b_first_frame_ = true, imu_need_init_ = true;
    int init_iter_num = 1;
    Vec3 mean_acc{0, 0, -1.0}, mean_gyr;
    Vec3 cov_acc{0.1, 0.1, 0.1}, cov_gyr{0.1, 0.1, 0.1};
    Vec3 angvel_last, acc_s_last;
    Mat3 Lidar_R_wrt_IMU = Mat3::identity();
    Vec3 Lidar_T_wrt_IMU;
    ImuSample last_imu_;
    double last_observation_end_time_ = 0.0;  // IMU_Processing.hpp:87 never initialises it; fresh heap reads as 0
};

// dense n x n inverse, LU with partial pivoting (Eigen's .inverse() for n > 4)
bool invert(const double *A, int n, double *out);
bool solve_first_columns(const double *A, int n, int m, double *X);

class LaserMapping {
   public:
    explicit LaserMapping(const dlt_lio_config &cfg);
    ~LaserMapping();
    bool ok() const { return dev_ != nullptr; }
    int create_rc() const { return create_rc_; }
    // pts_on_device: pts48 is a device pointer and observation_end_time is given by the caller
    int process_scan(const void *pts48, int n, double lidar_beg_time, const ImuSample *imu, int n_imu, const dlt_lio_thermal *th,
                     dlt_lio_scan_out *out, bool pts_on_device = false, double observation_end_time_in = 0.0);
    void on_lidar_msg();       // feat_points_cbk, laserMapping.cpp:424-446
    void on_edge_count(int n); // tn_cbk, :491-498
    static int reduce_trampoline(void *self, double *result_dev, int n);  // forwards to reduce_fn

    dlt_lio_config cfg;
    dlt_handle dev_ = nullptr;
    dlt_lio_reduce_fn reduce_fn = nullptr;  // sharded map: sums the partial normal equations over the ranks
    void *reduce_ctx = nullptr;
    double *reduce_buf_dev = nullptr;
    bool insert_pending = false;  // dlt_map_incremental_async is in flight: its counts have not been adopted yet
    int last_added_ds = 0, last_added_raw = 0;
    int collect_insert(int *n_ds, int *n_raw);
    bool peers = false;  // dlt_lio_peer_attach: the sums over the ranks happen inside the kernels (peer mailboxes), no callback
    ImuProcess imu_;
    StatesGroup state, last_nodegared_state, last_state;
    std::vector<dlt_lio_iter> iters;
    std::string err;
    bool EKF_stop_flg = false, flg_EKF_inited = false, map_built = false, Localmap_Initialized = false;
    int dynamic_effect_featurepoints_threshold = 100;  // laserMapping.cpp:97
    long lidar_frame_counter_num = 0, lidar_cnt = 0;
    std::deque<int> effct_q;  // effct_feat_numQueue, laserMapping.cpp:192-193
    float LocalMap_min[3] = {0, 0, 0}, LocalMap_max[3] = {0, 0, 0};

   private:
    int fov_segment(const Vec3 &pos_LiD, int *deleted);  // lasermap_fov_segment, :313-369
    void zeta_blend(int effct_feat_num, const StatesGroup &state_propagat, const dlt_lio_thermal *th);  // :1105-1131
    dlt_iekf_block *iekf_blk_ = nullptr;  // host copy of the device-resident loop's block
    bool flg_first_scan = true;
    double first_lidar_time = 0.0;
    double zeta_l = 0.0, zeta_t = 0.0;
    int create_rc_ = 0;
};

}  // namespace dlt_host

struct dlt_lio_s {
    dlt_host::LaserMapping *lm;
};
