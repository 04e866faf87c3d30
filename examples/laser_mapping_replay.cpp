// examples/laser_mapping_replay.cpp
//
// The reference-side binding of INTEGRATION.md section 2 as a program that compiles and runs without ROS: the body of
// `while (sync_packages(Measures))` in eskf_lio/src/laserMapping.cpp:731-1177 replaced by ONE call into libdaliti_b200.so
// (include/daliti_b200_lio.h).  Everything ROS would deliver -- the /laser_cloud_surf cloud (48-byte PointXYZINormal
// records), its stamp, the IMU samples sync_packages selected -- is synthesised here: a sensor gliding through a room.
// It is an integration example and a link/ABI check (tests/test_example.py builds it against the kernel-logic emulator
// on CPU; against the CUDA library it needs a B200).  Plain C++11 against a plain C header: no torch, no Python.
//
//   g++ -std=c++14 -O2 -I include examples/laser_mapping_replay.cpp -L daliti_b200/lib -ldaliti_b200 -o replay
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "daliti_b200_lio.h"

namespace {
struct PointXYZINormal {  // pcl::PointXYZINormal as eskf_lio fills it (my_utility.h:57, feature_extract.cpp:337-346)
    float x, y, z, pad0;
    float normal_x /* time ratio */, normal_y /* ring */, normal_z /* sweep span, s */, pad1;
    float intensity, curvature, pad2, pad3;
};
static_assert(sizeof(PointXYZINormal) == 48, "wire format");

struct MeasureGroup {  // common_lib.h MeasureGroup without the ROS types
    double lidar_beg_time;
    std::vector<PointXYZINormal> lidar;
    std::vector<double> imu7;  // t, acc xyz, gyr xyz per sample
};

const double kSweep = 0.1, kG = 9.81;
const double kRoom[6] = {-12.0, -9.0, 0.0, 14.0, 8.0, 6.0};  // xmin ymin zmin xmax ymax zmax
const double kVel[3] = {0.8, 0.3, 0.0}, kStart[3] = {-4.0, -2.0, 1.6};

void sensor_at(double t, double p[3]) {
    for (int i = 0; i < 3; i++) p[i] = kStart[i] + kVel[i] * t;
}

// what sync_packages would hand over for sweep k: 16 beams x 720 azimuths against the room's inner faces
MeasureGroup synthesize(int k) {
    MeasureGroup m;
    m.lidar_beg_time = 100.0 + k * kSweep;
    const int beams = 16, az = 720;
    unsigned rng = 12345u + 977u * (unsigned)k;
    for (int a = 0; a < az; a++) {
        const double ratio = (double)a / az, t = k * kSweep + ratio * kSweep;
        double p[3];
        sensor_at(t, p);
        for (int b = 0; b < beams; b++) {
            const double el = (-15.0 + 30.0 * b / (beams - 1)) * M_PI / 180.0, th = 2.0 * M_PI * ratio;
            const double d[3] = {std::cos(el) * std::cos(th), std::cos(el) * std::sin(th), std::sin(el)};
            double range = 1e9;
            for (int ax = 0; ax < 3; ax++) {
                if (d[ax] > 1e-9) range = std::fmin(range, (kRoom[3 + ax] - p[ax]) / d[ax]);
                if (d[ax] < -1e-9) range = std::fmin(range, (kRoom[ax] - p[ax]) / d[ax]);
            }
            rng = rng * 1664525u + 1013904223u;
            range += 0.01 * (((rng >> 8) & 0xFFFF) / 32768.0 - 1.0);  // +-1 cm range noise
            PointXYZINormal q;
            std::memset(&q, 0, sizeof(q));
            q.x = (float)(range * d[0]);  // body frame at the moment of the shot: deskew has real work to do
            q.y = (float)(range * d[1]);
            q.z = (float)(range * d[2]);
            q.normal_x = (float)ratio;
            q.normal_y = (float)b;
            q.normal_z = (float)kSweep;
            q.intensity = 10.f + b;
            m.lidar.push_back(q);
        }
    }
    for (int i = 0; i < 20; i++) {  // 200 Hz, body at constant velocity: specific force = -gravity
        const double t = m.lidar_beg_time + (i + 0.5) * kSweep / 20.0;
        const double s[7] = {t, 0.0, 0.0, kG, 0.0, 0.0, 0.0};
        m.imu7.insert(m.imu7.end(), s, s + 7);
    }
    return m;
}
}  // namespace

int main(int argc, char **argv) {
    const int n_scans = argc > 1 ? std::atoi(argv[1]) : 8;
    // ---- main(), after the rosparam block (laserMapping.cpp:651-667)
    dlt_lio_config cfg;
    dlt_lio_default_config(&cfg);
    cfg.max_iteration = 4;         // mapping/max_iteration
    cfg.dev.ds_scan = 0.5f;        // mapping/filter_size_surf
    cfg.dev.ds_map = 0.5f;         // mapping/filter_size_map
    cfg.dev.max_scan_points = 16384;
    cfg.dev.max_map_points = 1 << 16;
    cfg.featptsThreshold = 30;     // common/featptsThreshold
    dlt_lio g_lio = nullptr;
    int rc = dlt_lio_create(&cfg, &g_lio);
    if (rc != DLT_OK) {
        std::fprintf(stderr, "dlt_lio_create failed (%d): daliti_b200 has no CPU path, a CUDA device is required\n", rc);
        return 2;
    }
    // IMU_Initial's ~100 stationary samples (IMU_Processing.hpp:385-405) are skipped with a known mean
    const double mean_acc[3] = {0.0, 0.0, kG}, last_imu[7] = {100.0 - 0.0025, 0.0, 0.0, kG, 0.0, 0.0, 0.0};
    dlt_lio_force_imu_ready(g_lio, mean_acc, last_imu);
    std::vector<double> s(612, 0.0);
    s[0] = s[4] = s[8] = 1.0;       // rot_end
    s[12] = s[16] = s[20] = 1.0;    // R_L_I
    for (int i = 0; i < 3; i++) s[9 + i] = kStart[i], s[24 + i] = kVel[i];
    s[35] = -kG;
    for (int i = 0; i < 24; i++) s[36 + 25 * i] = 1.0;  // cov = I, as the StatesGroup constructor leaves it
    dlt_lio_set_state(g_lio, s.data(), 1);

    double worst = 0.0;
    int updates = 0;
    // ---- while (sync_packages(Measures)) { ... }   laserMapping.cpp:731-1177
    for (int k = 0; k < n_scans; k++) {
        MeasureGroup Measures = synthesize(k);
        dlt_lio_on_lidar_msg(g_lio);  // feat_points_cbk bookkeeping (:424-446)
        dlt_lio_thermal th;
        std::memset(&th, 0, sizeof(th));  // thermal odometry offline
        th.delta_quat[0] = th.l2l_quat[0] = 1.0;
        dlt_lio_scan_out out;
        rc = dlt_lio_process_scan(g_lio, Measures.lidar.data(), (int)Measures.lidar.size(), Measures.lidar_beg_time, Measures.imu7.data(),
                                  (int)Measures.imu7.size() / 7, &th, &out);
        if (rc != DLT_OK) {
            std::fprintf(stderr, "scan %d: %s\n", k, dlt_lio_last_error(g_lio));
            return 1;
        }
        if (!out.had_points) continue;  // "No point, not ready for odometry, skip this scan" (:755-759)
        if (out.added < 0) {  // async_insert (default): map_incremental is still running; a logger that wants its counts waits here
            dlt_lio_collect_insert(g_lio, &out.n_added_ds, &out.n_added_raw);
            out.added = out.n_added_ds + out.n_added_raw;
        }
        dlt_lio_get_state(g_lio, s.data());
        double truth[3];
        sensor_at((k + 1) * kSweep, truth);
        const double err = std::sqrt((s[9] - truth[0]) * (s[9] - truth[0]) + (s[10] - truth[1]) * (s[10] - truth[1]) + (s[11] - truth[2]) * (s[11] - truth[2]));
        std::printf("scan %d: raw %d down %d map %d iters %d added %d ekf_stop %d degenerate %d eig_min %.3g | pos %.4f %.4f %.4f err %.4f m\n", k,
                    out.n_raw, out.n_down, out.map_points_before, out.n_iters, out.added, out.ekf_stop, out.degenerate, out.eigvals[0], s[9], s[10],
                    s[11], err);
        if (out.did_update) {
            updates++;
            worst = std::fmax(worst, err);
            // odomAftMapped.pose.covariance[0] = out.degenerate;   the slot imuPreintegration.cpp:290 reads
        }
    }
    dlt_lio_destroy(g_lio);
    std::printf("updates %d worst position error %.4f m\n", updates, worst);
    return (updates >= n_scans - 2 && worst < 0.05) ? 0 : 1;
}
