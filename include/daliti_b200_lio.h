/* include/daliti_b200_lio.h -- C ABI of the host-side per-scan update.
 *
 * Mirrors the body of eskf_lio's main loop, `while (sync_packages(Measures)) { ... }`
 * (eskf_lio/src/laserMapping.cpp:731-1177), with the device path of daliti_b200.h
 * underneath: IMU forward propagation and the 24-state Kalman algebra stay on the host
 * (they are sequential and tiny), everything per-point runs on the B200.
 *
 * What the ROS node keeps: callbacks, sync_packages, publishing (laserMapping.cpp:424-574,
 * 1183-1290).  What it hands over, per scan: the PointXYZINormal cloud of
 * /laser_cloud_surf, its header stamp, the IMU samples sync_packages selected, and the
 * scalars the thermal callbacks leave behind.  What it gets back: the updated StatesGroup,
 * EKF_stop_flg, effct_feat_num and the degeneracy eigen-values.
 */
#ifndef DALITI_B200_LIO_H
#define DALITI_B200_LIO_H

#include "daliti_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dlt_lio_s *dlt_lio;

typedef struct dlt_lio_config {
    dlt_config dev;          /* device path configuration                                         */
    int max_iteration;       /* mapping/max_iteration            laserMapping.cpp:656              */
    double cube_len;         /* mapping/cube_side_length         laserMapping.cpp:659              */
    int featptsThreshold;    /* common/featptsThreshold          laserMapping.cpp:654              */
    double beta;             /* common/beta                      laserMapping.cpp:664              */
    float det_range;         /* DET_RANGE                        laserMapping.cpp:304              */
    double extrinT[3];       /* mapping/extrinsic_T              laserMapping.cpp:661              */
    double extrinR[9];       /* mapping/extrinsic_R (row-major)  laserMapping.cpp:662              */
    double degeneracy_eig_threshold; /* new: flag when min eigenvalue of HtH[0:6,0:6] is below     */
    int device_loop;         /* where the iteration loop :820-1102 runs.
                                0: on the host, one dlt_measure round trip per iteration (Kalman algebra as the
                                   reference writes it);
                                2: resident on the device (dlt_iekf_update), zeta blend + map_incremental driven
                                   by the host: two synchronisations per scan;
                                1: blend and insert also queued on the device behind the loop: one synchronisation
                                   (on a sharded map this needs attached peers, dlt_lio_peer_attach; otherwise it acts as 2);
                               -1 (default): 0 on a single-GPU map, 2 on a sharded map (measured)      */
    int async_insert;        /* 1 (default): map_incremental (:582-630, 1164-1168) runs off the critical path, on its own
                                stream behind the last evaluation of the measurement model: dlt_lio_process_scan returns
                                once the pose is final, with added / n_added_* = -1; the counts (and an overflow of the
                                map) are reported by dlt_lio_collect_insert or picked up by the next scan.  Unsharded maps and
                                sharded maps with attached peers; 0 = wait for the insert inside the call, as the
                                reference does.                                                                        */
} dlt_lio_config;

/* State left behind by tis_cbk / tn_cbk (laserMapping.cpp:471-498): g_tis_odom_delta and
 * g_tis_odom_delta_lframe2lframe as position, quaternion (w x y z), linear velocity and
 * the pose.covariance[0..7] slots odomToStateGruop reads (laserMapping.cpp:218-239).           */
typedef struct dlt_lio_thermal {
    int tis_online, recv_n;
    double delta_pos[3], delta_quat[4], delta_vel[3], cov_slots[8];
    double l2l_pos[3], l2l_quat[4], l2l_vel[3], l2l_cov_slots[8];
} dlt_lio_thermal;

/* One row of Log/mat_out.txt (laserMapping.cpp:936-937) plus the algebra of the iteration.    */
typedef struct dlt_lio_iter {
    int iter, effct_feat_num, converged, ekf_stop, did_match, n_down;
    double total_residual, res_mean_last;
    double HtH[144], Htr[12], pose_in[24], state_out[36], solution[24];
} dlt_lio_iter;

typedef struct dlt_lio_scan_out {
    int had_points, built_map, did_update, ekf_stop;
    int n_raw, n_down, map_points_before, deleted, added, n_iters;
    int n_added_ds, n_added_raw;
    double eigvals[6], eigvecs[36];  /* of the last iteration's HtH[0:6,0:6]                      */
    double state_prop[36];           /* state_propagat                     laserMapping.cpp:752    */
    int degenerate;                  /* eigvals[0] < degeneracy_eig_threshold; slot odom.pose.covariance[0]
                                        read at daliti/src/lidar_odometry/imuPreintegration.cpp:290 */
    int reserved;
    double t_deskew, t_voxel, t_iterate, t_insert, t_delete, t_total; /* host wall clock, seconds  */
} dlt_lio_scan_out;

void dlt_lio_default_config(dlt_lio_config *cfg);
int dlt_lio_create(const dlt_lio_config *cfg, dlt_lio *out);
int dlt_lio_destroy(dlt_lio h);
const char *dlt_lio_last_error(dlt_lio h);
dlt_handle dlt_lio_device(dlt_lio h);            /* the device handle (map access, exports)        */

/* feat_points_cbk / tn_cbk bookkeeping               laserMapping.cpp:424-446, 491-498            */
int dlt_lio_on_lidar_msg(dlt_lio h);
int dlt_lio_on_edge_count(dlt_lio h, int recv_n);
/* skip ImuProcess::IMU_Initial's ~100 samples (IMU_Processing.hpp:385-405) with a known mean_acc
 * and last IMU sample (t, acc[3], gyr[3])                                                        */
int dlt_lio_force_imu_ready(dlt_lio h, const double *mean_acc3, const double *last_imu7);
/* StatesGroup as 36 + 576 doubles: rot_end[9] pos_end[3] R_L_I[9] T_L_I[3] vel_end[3] bias_g[3]
 * bias_a[3] gravity[3] cov[576]                       common_lib.h:73-228                         */
int dlt_lio_get_state(dlt_lio h, double *state612);
int dlt_lio_set_state(dlt_lio h, const double *state612, int also_last_states);
int dlt_lio_get_flags(dlt_lio h, int *flags8);   /* EKF_stop_flg, flg_EKF_inited, threshold, lidar_cnt, localmap_init, queue size, map_built, imu_ready */
int dlt_lio_get_localmap(dlt_lio h, float *box6);

/* The per-scan update.  pts48: n PointXYZINormal records; imu7: n_imu rows of t, acc[3], gyr[3]. */
int dlt_lio_process_scan(dlt_lio h, const void *pts48, int n, double lidar_beg_time, const double *imu7, int n_imu,
                         const dlt_lio_thermal *thermal, dlt_lio_scan_out *out);
/* feat_points_cbk (:424-446) may hand a scan over as soon as it arrives: its upload then overlaps the update of the previous
 * scan (dlt_scan_prefetch); the dlt_lio_process_scan call for the same buffer finds the records already on the device.  */
int dlt_lio_prefetch_scan(dlt_lio h, const void *pts48, int n);
/* Same with the scan already resident in DEVICE memory (pts48_dev); observation_end_time is then
 * passed explicitly (the host cannot read points.back().normal_z, laserMapping.cpp:546).            */
int dlt_lio_process_scan_dev(dlt_lio h, const void *pts48_dev, int n, double lidar_beg_time, double observation_end_time,
                             const double *imu7, int n_imu, const dlt_lio_thermal *thermal, dlt_lio_scan_out *out);
/* The same starting one node earlier: the sensor's PointCloud2 payload (HOST memory) goes through the
 * front end of feature_extract.cpp:264-450 (feature_enabled = 0; dlt_frontend_sample) on the device and straight into the
 * update -- the /laser_cloud_surf hop and its 48-byte AoS round trip disappear.  header_stamp = cloudHeader.stamp (:268);
 * n_sampled (optional) <- sampleCloud->size().                                                                          */
int dlt_lio_process_cloud(dlt_lio h, const void *cloud_data, int n_points, const dlt_cloud_layout *layout, int sensor, int point_filter_num,
                          float lidar_min_range, float lidar_max_range, double header_stamp, const double *imu7, int n_imu,
                          const dlt_lio_thermal *thermal, dlt_lio_scan_out *out, int *n_sampled);
/* Sharded map (dev.shard_count > 1): after every evaluation of the measurement model the partial normal
 * equations (n = 158 doubles at result_dev, DEVICE memory owned by the caller, 256 doubles) are handed to
 * `reduce`, which must sum them over the ranks in place on the handle's stream (e.g. ncclAllReduce /
 * torch.distributed.all_reduce) and return 0.  With device_loop it is called max_iteration times per scan
 * on every rank and must only ENQUEUE the reduction (no host wait).  The same callback then sums the per-point
 * map_incremental decisions (n = feats_down_size doubles at the pointer it is given -- it must reduce THAT buffer,
 * not one of its own): the owner of a query decides, every rank inserts into its tiles + halo.                     */
typedef int (*dlt_lio_reduce_fn)(void *ctx, double *result_dev, int n);
int dlt_lio_set_reduce(dlt_lio h, dlt_lio_reduce_fn reduce, void *ctx, double *result_dev /* NULL: the handle's own buffer */);
/* The same sums over NVLink peer memory, from inside the kernels, instead of a callback (dlt_peer_export / dlt_peer_attach
 * in daliti_b200.h): every rank exports a DLT_PEER_BLOB_BYTES blob, the node gathers the blobs of all dev.shard_count ranks in
 * rank order and hands them to dlt_lio_peer_attach on every rank.  Afterwards no reduce callback is needed (one that is set
 * is ignored); the per-iteration sum costs no launch and no host involvement.                                           */
int dlt_lio_peer_export(dlt_lio h, unsigned char *blob);
int dlt_lio_peer_attach(dlt_lio h, const unsigned char *blobs);
int dlt_lio_peer_detach(dlt_lio h); /* back to the callback; a handle attaches at most once */
/* async_insert: wait for the map_incremental of the last scan and report its add counts (the reference logs them as
 * kdtree_incremental add_point_size, laserMapping.cpp:626-629).  Without a pending insert: the last counts again.  */
int dlt_lio_collect_insert(dlt_lio h, int *n_added_ds, int *n_added_raw);
int dlt_lio_get_iters(dlt_lio h, dlt_lio_iter *iters, int cap);

/* ---- many independent sequences on one GPU, driven natively (BASELINE config: a batch of independent sequences; the reference
 * runs one sequence per process and has no counterpart).  Every sequence has its own dlt_lio handle (own map, state, CUDA
 * streams).  n_threads worker threads (0 = one per host core, at most n_seqs) each replay several sequences as coroutines:
 * wherever one sequence's update waits for the device the thread goes on with the next sequence's.  Per scan the call is
 * exactly on_lidar_msg x lidar_msgs, [dlt_lio_prefetch_scan of scan k+1], dlt_lio_process_scan(_dev) of scan k; results are
 * those of the same calls made one by one.  Returns the first non-zero rc of any sequence (each is reported in seqs[i].rc).  */
typedef struct dlt_lio_seq_scan {
    const void *pts48;            /* n PointXYZINormal records, host (pinned for prefetch) or device memory              */
    int n, pts_on_device;
    double lidar_beg_time;
    double observation_end_time;  /* device input only (laserMapping.cpp:546)                                             */
    const double *imu7;           /* n_imu rows of t, acc[3], gyr[3]                                                      */
    int n_imu;
    int lidar_msgs;               /* dlt_lio_on_lidar_msg calls in front of this scan (feat_points_cbk, :424-446)         */
} dlt_lio_seq_scan;
typedef struct dlt_lio_seq {
    dlt_lio h;
    const dlt_lio_seq_scan *scans;
    int n_scans;
    int prefetch;                 /* host buffers: upload scan k+1 while scan k is processed                              */
    const dlt_lio_thermal *thermal; /* optional, applied to every scan                                                    */
    dlt_lio_scan_out *outs;       /* optional, n_scans entries                                                            */
    int rc, n_done;               /* out: status of this sequence, scans completed                                        */
} dlt_lio_seq;
int dlt_lio_replay_sequences(dlt_lio_seq *seqs, int n_seqs, int n_threads);
/* IMUpose list of the last scan's forward propagation (22 doubles each)                          */
int dlt_lio_get_imu_poses(dlt_lio h, double *pose22, int cap);

#ifdef __cplusplus
}
#endif
#endif /* DALITI_B200_LIO_H */
