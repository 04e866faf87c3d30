/* include/daliti_b200.h -- C ABI of the B200 scan-to-map measurement path.
 *
 * Drop-in boundary for DaLiTI's eskf_lio LiDAR measurement update.  The reference has no
 * function at this boundary: the update is inlined in main(),
 * eskf_lio/src/laserMapping.cpp:731-1177.  Each entry point below names the reference
 * statements it replaces (paths relative to the DaLiTI tree).  Plain pointers and sizes
 * only; all buffers are caller-owned HOST memory unless a name ends in `_dev`.
 *
 * Conventions: every function returns 0 on success or a DLT_E_* code; a handle owns one
 * CUDA stream and is not thread-safe; there is no CPU fallback -- dlt_create fails with
 * DLT_E_NO_DEVICE when no CUDA device is usable.
 *
 * Layouts
 *   xyzi   : float[4] per point  = x, y, z, intensity
 *   pts48  : pcl::PointXYZINormal, 48 bytes = x y z 1 | normal_x (time ratio) normal_y (ring)
 *            normal_z (sweep span, s) 0 | intensity curvature pad pad
 *            (eskf_lio/include/my_utility.h:57, eskf_lio/src/feature_extract.cpp:337-346)
 *   pose24 : double[24] = rot_end[9] (row-major), pos_end[3], R_L_I[9], T_L_I[3]
 *            (StatesGroup, eskf_lio/include/common_lib.h:219-222)
 *   imu pose: double[22] = offset_time, acc[3], gyr[3], vel[3], pos[3], rot[9]
 *            (eskf_lio/msg/Pose6D.msg)
 */
#ifndef DALITI_B200_H
#define DALITI_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define DLT_OK 0
#define DLT_E_INVALID 1      /* bad argument */
#define DLT_E_NO_DEVICE 2    /* no usable CUDA device (there is no CPU path) */
#define DLT_E_CUDA 3         /* CUDA runtime error, see dlt_last_error */
#define DLT_E_CAPACITY 4     /* a configured capacity was exceeded */
#define DLT_E_STATE 5        /* call sequence error (e.g. measure before a scan is set) */

typedef struct dlt_handle_s *dlt_handle;

typedef struct dlt_config {
    float ds_scan;        /* mapping/filter_size_surf, VoxelGrid leaf (feat.yaml:47)              */
    float ds_map;         /* mapping/filter_size_map, ikd-Tree downsample size (feat.yaml:48)      */
    float max_sq_dist;    /* match gate on the 5th neighbour, 5 (laserMapping.cpp:853)             */
    float plane_thr;      /* esti_plane threshold, 0.1 (laserMapping.cpp:863)                      */
    int extrinsic_est_en; /* mapping/extrinsic_est_en (laserMapping.cpp:968)                       */
    int device;           /* CUDA device ordinal                                                   */
    int max_scan_points;  /* capacity: raw points per scan                                         */
    int max_map_points;   /* capacity: live map points                                             */
    long long voxel_bitmap_bits; /* capacity of the VoxelGrid occupancy bitmap (0 = 2^27)          */
    int shard_rank;       /* spatial map sharding across GPUs: this rank ...                       */
    int shard_count;      /* ... of shard_count (1 = unsharded)                                    */
    int shard_tile_shift; /* tile edge = search-cell edge * 2^shift (0 = default 5)                */
} dlt_config;

/* Outputs of one evaluation of the measurement model.  What the reference later does with
 * Hsub / meas_vec / K factors through these (laserMapping.cpp:1015-1032, 1084):
 *   K z = K_1[:, :12] H^T r,  K H = K_1[:, :12] H^T H.                                          */
typedef struct dlt_measure_out {
    double HtH[144];        /* 12x12 row-major  H^T H   (laserMapping.cpp:1015)                    */
    double Htr[12];         /* H^T meas_vec     (meas_vec = -pd2, laserMapping.cpp:977)            */
    double total_residual;  /* sum of res_last over effective points (laserMapping.cpp:893)        */
    int effct_feat_num;     /* laserMapping.cpp:885-896                                            */
    int n_down;             /* feats_down_size                                                     */
    int n_unresolved;       /* queries whose exact neighbours are deferred to map_incremental      */
    int reserved;
} dlt_measure_out;

void dlt_default_config(dlt_config *cfg);
int dlt_create(const dlt_config *cfg, dlt_handle *out);
int dlt_destroy(dlt_handle h);
const char *dlt_last_error(dlt_handle h);
/* Adopt an external CUDA stream (cudaStream_t) for all work of this handle, or NULL to go
 * back to the handle's own stream.  Used to order work with torch / NCCL.                       */
int dlt_set_stream(dlt_handle h, void *cuda_stream);
/* The stream (cudaStream_t) the handle currently launches on: what a reduce callback has to enqueue its collective on.     */
void *dlt_stream(dlt_handle h);
int dlt_sync(dlt_handle h);

/* ---- map: replaces the global `KD_TREE<PointType> ikdtree` (laserMapping.cpp:164) ---------- */
/* ikdtree.Build(feats_down_world->points)                       laserMapping.cpp:790            */
int dlt_map_build(dlt_handle h, const float *xyzi, int n);
/* first-scan map initialisation: pointBodyToWorld over feats_down, then ikdtree.Build
 *                                                               laserMapping.cpp:780-791        */
int dlt_map_build_from_scan(dlt_handle h, const double *pose24);
/* ikdtree.Add_Points(points, downsample_on)                     laserMapping.cpp:627-628        */
int dlt_map_add(dlt_handle h, const float *xyzi, int n, int downsample_on);
/* ikdtree.Delete_Point_Boxes(cub_needrm), boxes = min xyz, max xyz   laserMapping.cpp:368       */
int dlt_map_delete_boxes(dlt_handle h, const float *boxes6, int n_boxes, int *deleted);
/* ikdtree.validnum()                                            laserMapping.cpp:794            */
int dlt_map_valid_count(dlt_handle h, int *n);
/* ikdtree.flatten(Root_Node, PCL_Storage, NOT_RECORD)           laserMapping.cpp:1171-1174      */
int dlt_map_export(dlt_handle h, float *xyzi, int cap, int *n);
/* ikdtree.Nearest_Search(point, 5, points_near, sq_dis) for nq world-frame points:
 * out_xyzi nq*5*4 floats (ascending), out_d2 nq*5, out_cnt nq    laserMapping.cpp:850           */
int dlt_map_knn(dlt_handle h, const float *q_xyz, int nq, float *out_xyzi, float *out_d2, int *out_cnt);

/* ---- scan preparation ---------------------------------------------------------------------- */
/* ImuProcess::UndistortPcl backward pass (IMU_Processing.hpp:332-370) over n_raw points with
 * the IMUpose list of the forward pass and the propagated end state; n_imu_pose < 2 copies
 * the points through.  Also the input of dlt_scan_downsample.                                   */
int dlt_scan_deskew(dlt_handle h, const void *pts48, int n_raw, const double *imu_pose22, int n_imu_pose, const double *pose24);
/* Optional double buffering: start uploading the NEXT scan's records (pinned HOST memory, untouched until consumed) on the
 * handle's copy stream now, e.g. from the message callback while the previous update is still running.  A later
 * dlt_scan_deskew / dlt_lio_process_scan with the same pointer and count uses that copy instead of uploading again.     */
int dlt_scan_prefetch(dlt_handle h, const void *pts48, int n_raw);
/* same with the PointXYZINormal records already resident in DEVICE memory (no host->device copy) */
int dlt_scan_deskew_dev(dlt_handle h, const void *pts48_dev, int n_raw, const double *imu_pose22, int n_imu_pose, const double *pose24);
/* downSizeFilterSurf.filter(*feats_down)                        laserMapping.cpp:775-776        */
int dlt_scan_downsample(dlt_handle h, int *n_down);
/* feats_undistort / feats_down read-back (publishing, laserMapping.cpp:1185,1208)               */
int dlt_scan_get_undistorted(dlt_handle h, float *xyzi, int cap, int *n);
int dlt_scan_get_down(dlt_handle h, float *xyzi, int cap, int *n);
/* set feats_down directly (body frame), bypassing deskew + VoxelGrid                            */
int dlt_scan_set_down(dlt_handle h, const float *xyzi, int n);
/* output slot of every raw point of the last dlt_scan_downsample (voxel assignment)             */
int dlt_scan_get_voxel_of_point(dlt_handle h, int *slot, int cap);

/* ---- LiDAR front end with feature_enabled = 0 --------------------------
 * FeatureExtract::cachePointCloud (eskf_lio/src/feature_extract.cpp:264-423) fused with samplePointCloud
 * (:425-450): the sensor's PointCloud2 records (HOST memory, `layout` = point_step and field offsets of the
 * message) become the PointXYZINormal records of /laser_cloud_surf -- normal_x = time ratio, normal_y = ring,
 * normal_z = sweep span; every point_filter_num-th record (feat.yaml:21), range gate (:445), order preserved --
 * directly in DEVICE memory, where dlt_scan_deskew_dev / dlt_lio_process_scan_dev read them.
 * Field types per sensor (eskf_lio/include/my_utility.h:19-54): Velodyne / Livox x y z intensity f32, ring u16,
 * time f32; Ouster the same with t u32 [ns]; RoboSense intensity u8, timestamp f64 (non-finite points dropped).   */
#define DLT_SENSOR_VELODYNE 0
#define DLT_SENSOR_LIVOX 1
#define DLT_SENSOR_OUSTER 2
#define DLT_SENSOR_ROBOSENSE 3
typedef struct dlt_cloud_layout {
    int point_step;                                  /* sensor_msgs/PointCloud2.point_step                       */
    int off_x, off_y, off_z, off_intensity, off_ring, off_time; /* PointField offsets                            */
} dlt_cloud_layout;
/* pts48_dev: the handle's record buffer (valid until the next scan call); n_out: sampleCloud->size();
 * timespan: what cachePointCloud leaves in `timespan` (timeScanEnd = stamp + timespan, :387);
 * sweep_span: normal_z of the records (laserMapping's observation_end_time = lidar_beg_time + points.back().normal_z);
 * stamp_shift: what it subtracts from the header stamp (RoboSense, :383), else 0.                                  */
int dlt_frontend_sample(dlt_handle h, const void *cloud_data, int n_points, const dlt_cloud_layout *layout, int sensor, int point_filter_num,
                        float lidar_min_range, float lidar_max_range, void **pts48_dev, int *n_out, double *timespan, double *sweep_span,
                        double *stamp_shift);

/* the first n records of the last dlt_frontend_sample, back on the host (publishing /laser_cloud_surf, tests)        */
int dlt_frontend_read(dlt_handle h, void *pts48, int n);

/* ---- the measurement model, one IEKF iteration ---------------------------------------------- */
/* laserMapping.cpp:829-979 at the given state: transform, (re)match when do_match
 * (iterCount == 0 || rematch_en, :847), plane fit, residual + gates, Jacobian, and the
 * reduction to H^T H / H^T r.  Nearest_Points, point_selected_surf and the cached planes
 * persist in the handle between calls of one scan.                                              */
int dlt_measure(dlt_handle h, const double *pose24, int do_match, dlt_measure_out *out);
/* Same, but leaves the result block (HtH[144] Htr[12] count res_sum = 158 doubles; the buffer must hold 256)
 * in DEVICE memory at result_dev without synchronising -- the partial sums of one map shard,
 * ready for an NCCL all-reduce of the 158 doubles.                                              */
int dlt_measure_dev(dlt_handle h, const double *pose24, int do_match, double *result_dev);
/* The handle's own 256-double result buffer in DEVICE memory (a valid result_dev for the calls around it).        */
double *dlt_result_dev(dlt_handle h);
/* Read a result block produced by dlt_measure_dev (possibly all-reduced in between) back to the host. */
int dlt_fetch_result(dlt_handle h, const double *result_dev, dlt_measure_out *out);
/* laserCloudOri / coeffSel of the last dlt_measure (published as /cloud_effected,
 * laserMapping.cpp:891-892, 1213-1227): body-frame xyzi and (normal, pd2) per effective point   */
int dlt_effective_points(dlt_handle h, float *xyzi, float *coeff, int cap, int *n);
/* Nearest_Points of the last match pass: nbr n_down*5*4 floats (x y z d2), cnt n_down,
 * selected n_down (point_selected_surf after the last dlt_measure)                              */
int dlt_get_nearest(dlt_handle h, float *nbr, int *cnt, unsigned char *selected, int cap);
/* Degradation output (new; the reference has no eigen check): eigen-decomposition
 * of HtH[0:6,0:6] of the last dlt_measure, computed on the device (parallel Jacobi, one warp)
 * when asked for.  Ascending eigenvalues; eigenvectors in the columns of the row-major 6x6.      */
int dlt_degeneracy(dlt_handle h, double *eigvals6, double *eigvecs36);
/* Enqueue that computation and its device->host copy on the handle's side stream without waiting: it
 * overlaps whatever is queued next on the main stream (dlt_map_incremental); a later dlt_degeneracy
 * then only synchronises.                                                                          */
int dlt_degeneracy_begin(dlt_handle h);

/* ---- the iteration loop resident on the device ------------------------
 * laserMapping.cpp:820-1102 without a host round trip between iterations: per iteration the
 * measurement model above, the degradation window (:899-918), the Kalman update (:1012-1053, with
 * the gain K_1 = (H^T H + (P/R)^-1)^-1 evaluated in the algebraically equal covariance form
 * P' - P' U (I + H P'_11)^-1 H U^T P', P' = P/R: one 6 x 6 or 12 x 12 solve, no inversions) or
 * the stop branch (:1054-1063), the rematch / convergence control (:1069-1101) and, at loop
 * exit, the covariance update (:1084-1085).  The caller fills the `in` and `in / out` parts of
 * one block; dlt_iekf_update copies it to the device, enqueues max_iteration rounds of
 * {match pass, residual pass, solve} whose kernels turn into no-ops once the loop has ended,
 * and reads the block back with ONE synchronisation.  State layout = StatesGroup flat:
 * rot_end[9] pos_end[3] R_L_I[9] T_L_I[3] vel_end[3] bias_g[3] bias_a[3] gravity[3] (+ cov[576]). */
#define DLT_IEKF_MAX_ITER 16
typedef struct dlt_iekf_iter {          /* one row of Log/mat_out.txt (:936-937) + the algebra  */
    int iter, effct_feat_num, converged, ekf_stop, did_match, reserved;
    double total_residual;
    double HtH[144], Htr[12], pose_in[24], state_out[36], solution[24];
} dlt_iekf_iter;
typedef struct dlt_iekf_block {
    /* in */
    double state_propagat[36];          /* :752                                                  */
    double thermal_delta[36];           /* odomToStateGruop(g_tis_odom_delta), :1057, flat state */
    double laser_point_cov;             /* LASER_POINT_COV, :76 (the gain uses state.cov / this) */
    int threshold;                      /* dynamic_effect_featurepoints_threshold, :908          */
    int max_iteration;                  /* NUM_MAX_ITERATIONS (<= DLT_IEKF_MAX_ITER)             */
    int finish;                         /* 1: also run the zeta blend (:1105-1131) and map_incremental (:1164-1168)
                                           on the device behind the loop, before the one synchronisation; cleared
                                           on return when that was not possible (sharded map without attached peers) */
    int blend_mode;                     /* which branch of :1107 applies: 1 = LiDAR/IMU blend, 2 = thermal  */
    double last_state[36];              /* last_state, :1131 (finish only)                        */
    double l2l_state[36];               /* odomToStateGruop(g_tis_odom_delta_lframe2lframe), :1125 */
    double zeta_t, beta;                /* :1112, feat.yaml:8                                     */
    int lidar_cnt_lt_100;               /* :1113                                                  */
    int far_enqueued;                   /* set by dlt_iekf_update: the exact-neighbour fallback kernels are queued behind the loop */
    /* in / out, updated in place */
    double state[612];                  /* state at loop entry -> state at loop exit             */
    double last_nodegared[612];         /* last_nodegared_state, :1050                           */
    int effct_queue[10];                /* effct_feat_numQueue, oldest first, :899-905           */
    int queue_len, flg_EKF_inited;
    /* out */
    int n_iters, converged, ekf_stop, have_gain;
    int status;                         /* 0 ok, 1 = H^T H + (P/R)^-1 singular                   */
    int n_down, n_unresolved;
    int reserved1;                      /* VoxelGrid status of the scan (2 = capacity exceeded -> DLT_E_CAPACITY) */
    /* loop control, owned by the device (zeroed by dlt_iekf_update)                              */
    int iter, rematch_num, rematch_en, done;
    double zeta_l;                      /* in / out: :1110 (finish only)                          */
    double blend_state[36];             /* state after the zeta blend (its covariance is last_state's, :1119) */
    int insert_status;                  /* 0 not run (EKF stop / sharded / !finish), 1 done, 2 not done: more unresolved
                                           queries than one fallback chunk -- call dlt_map_incremental instead */
    int n_added_ds, n_added_raw;        /* PointToAdd / PointNoNeedDownsample sizes, :627-628     */
    int map_counters[5];                /* buckets, live points, map error, -, -                   */
    dlt_iekf_iter iters[DLT_IEKF_MAX_ITER];
} dlt_iekf_block;
/* downSizeFilterSurf.filter enqueued without waiting: feats_down_size stays on the device and is
 * reported by the next dlt_iekf_update (n_down).                                                 */
int dlt_scan_downsample_async(dlt_handle h);
/* reduce/reduce_ctx (optional): called once per enqueued iteration between the residual pass and the solve
 * with the 158-double partial normal equations in DEVICE memory; it must sum them over the ranks
 * in place on the handle's stream without blocking the host (sharded map, NCCL all-reduce).      */
typedef int (*dlt_reduce_fn)(void *ctx, double *result_dev, int n);
/* result_dev: where the residual pass leaves the 158 doubles handed to `reduce` (DEVICE memory, 256
 * doubles, caller-owned); NULL = the handle's own buffer.                                         */
int dlt_iekf_update(dlt_handle h, dlt_iekf_block *blk, dlt_reduce_fn reduce, void *reduce_ctx, double *result_dev);

/* ---- map_incremental()                                       laserMapping.cpp:582-630, 1167 - */
int dlt_map_incremental(dlt_handle h, const double *pose24, int flg_EKF_inited, int *n_add_downsample, int *n_add_raw);
/* The same, off the critical path: the kernels are queued on the handle's insert stream behind what the main stream holds
 * now and the call returns at once.  Every later call that reads or changes the map (the next match pass included) is
 * ordered behind them on the device; the next VoxelGrid only behind the classification pass that still reads the
 * downsampled scan.  dlt_map_incremental_collect waits on the host and reports the add counts and a map overflow.
 * Falls back to the synchronous form on a sharded map WITHOUT attached peers (the reduce callback runs on the caller's
 * stream) or when unresolved queries need the exact-neighbour fallback.                                             */
int dlt_map_incremental_async(dlt_handle h, const double *pose24, int flg_EKF_inited);
int dlt_map_incremental_collect(dlt_handle h, int *n_add_downsample, int *n_add_raw);
/* Sharded map (shard_count > 1): the rank that owns a query point decides whether it is added (it holds the
 * point's neighbours); `reduce` sums the n per-point decision codes (doubles in DEVICE memory) over the ranks in
 * place on the handle's stream, after which every rank inserts the accepted points that fall into its tiles + halo.
 * Unresolved far points are classified against this rank's tiles + halo only.                               */
int dlt_set_shard_reduce(dlt_handle h, dlt_reduce_fn reduce, void *ctx);

/* ---- sharded map: exchange over NVLink peer memory instead of the callbacks above (no reference counterpart) -- */
/* Every rank owns a mailbox in its HBM; peers store their partial normal equations / map_incremental decisions
 * straight into it from inside k_residual / k_incr_push and the receiving kernels add them up in rank order
 * (daliti_b200/csrc/dlt_peer.cuh), so the per-iteration all-reduce costs no launch and no library call.
 *   1. every rank: dlt_peer_export -> a DLT_PEER_BLOB_BYTES blob (CUDA IPC handle of the mailbox)
 *   2. the application gathers the blobs of all shard_count ranks in rank order (MPI, torch.distributed, a file ...)
 *   3. every rank: dlt_peer_attach(blobs).  From then on dlt_iekf_update ignores its reduce callback, dlt_measure /
 *      dlt_measure_dev return sums over the ranks, and dlt_map_incremental needs no dlt_set_shard_reduce.
 * Every rank must make the same sequence of these calls (they are collective).  A peer that never shows up makes
 * the waiting kernel give up after 5 s and the call return DLT_E_STATE; it does not hang the device.
 * Ranks may be processes (one per GPU) or handles on different devices of one process.  shard_count <= DLT_MAX_PEERS. */
#define DLT_MAX_PEERS 8
#define DLT_PEER_BLOB_BYTES 128
int dlt_peer_export(dlt_handle h, unsigned char *blob /* DLT_PEER_BLOB_BYTES */);
int dlt_peer_attach(dlt_handle h, const unsigned char *blobs /* shard_count x DLT_PEER_BLOB_BYTES, rank order */);
int dlt_peer_detach(dlt_handle h);

/* ---- instrumentation (no reference counterpart) ---------------------------------------------- */
/* Per-kernel-group device time from CUDA events on the launching stream.  Groups: 0 k_knn,
 * 1 k_residual, 2 deskew, 3 VoxelGrid, 4 map insert, 5 exact-neighbour fallback, 6 k_iekf_step, 7 k_knn8 alone (inside group 0).      */
/* Instrumentation: the handle's 16 device counters ([0] buckets allocated, [1] live points, [2] sticky map error, [5] unresolved
 * queries of the last match pass, [6] / [7] adds of the last map_incremental, [12] / [13] queries seen / searched again by the
 * rematch passes that reuse proven neighbour sets -- cumulative).  Synchronises the handle.                               */
int dlt_debug_counters(dlt_handle h, int *out16);
/* While THIS thread waits for the device inside any dlt_* call (result block of dlt_measure, stream / event waits) the library
 * calls fn(ctx) in a loop instead of spinning or blocking; NULL restores the default.  The multi-sequence driver
 * (dlt_lio_replay_sequences) uses it to run another sequence's update on the same thread meanwhile.                       */
int dlt_set_thread_wait_hook(void (*fn)(void *), void *ctx);
int dlt_set_profiling(dlt_handle h, int on);
int dlt_get_profile(dlt_handle h, double *ms8, long long *count8, int reset);
/* Spans recorded since the last dlt_get_profile(reset): (group, start ms, end ms) triples relative to the
 * first span, device time from the CUDA events (group 6 = k_iekf_step, 8 = k_eigen6 on the side stream). */
int dlt_get_timeline(dlt_handle h, double *kind_start_end, int cap, int *n);
/* SM clock stamps of the stages of k_iekf_step in the last dlt_iekf_update: n_iter x 16 values (kernel tuning)  */
int dlt_get_iekf_clocks(dlt_handle h, long long *clocks, int n_iter);
/* kernels launched by this library in this process so far                                         */
unsigned long long dlt_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DALITI_B200_H */
