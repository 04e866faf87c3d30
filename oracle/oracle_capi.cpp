// oracle/oracle_capi.cpp -- TEST INFRASTRUCTURE ONLY.
// extern "C" surface of the CPU oracle for ctypes (tests/, smoke(), bench.py cpu legs).
#include <dlfcn.h>

#include <string>

#include "oracle.hpp"

using namespace orc;

static RefApi g_ref;
static void *g_ref_dl = nullptr;

extern "C" {

// ---------------------------------------------------------------- reference library hookup
int orc_load_ref(const char *path) {
    if (g_ref.ok()) return 1;
    g_ref_dl = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!g_ref_dl) return 0;
#define SYM(field, name) g_ref.field = (decltype(g_ref.field))dlsym(g_ref_dl, name)
    SYM(create, "ikdref_create");
    SYM(destroy, "ikdref_destroy");
    SYM(build, "ikdref_build");
    SYM(knn, "ikdref_knn");
    SYM(add, "ikdref_add");
    SYM(delete_boxes, "ikdref_delete_boxes");
    SYM(validnum, "ikdref_validnum");
    SYM(size, "ikdref_size");
    SYM(has_root, "ikdref_has_root");
    SYM(flatten, "ikdref_flatten");
#undef SYM
    return g_ref.ok() ? 1 : 0;
}
int orc_ref_available() { return g_ref.ok() ? 1 : 0; }

// ---------------------------------------------------------------- small kernels
void orc_esti_plane_batch(const float *pts, int n, float thr, float *pabcd, uint8_t *ok) {
    for (int i = 0; i < n; i++) {
        float p[5][3];
        std::memcpy(p, pts + (size_t)i * 15, sizeof(p));
        ok[i] = esti_plane(pabcd + 4 * (size_t)i, p, thr) ? 1 : 0;
    }
}

int orc_voxel_grid(const float *pts48, int n, float leaf, int stable, float *out48, int cap, int *voxel_of_point,
                   unsigned *idx_of_point) {
    std::vector<Pt> out;
    std::vector<int> vop;
    std::vector<unsigned> iop;
    int m = voxel_grid(reinterpret_cast<const Pt *>(pts48), n, leaf, out, stable != 0, voxel_of_point ? &vop : nullptr,
                       idx_of_point ? &iop : nullptr);
    for (int i = 0; i < m && i < cap; i++) std::memcpy(out48 + 12 * (size_t)i, &out[i], 48);
    if (voxel_of_point)
        for (size_t i = 0; i < vop.size(); i++) voxel_of_point[i] = vop[i];
    if (idx_of_point)
        for (size_t i = 0; i < iop.size(); i++) idx_of_point[i] = iop[i];
    return m;
}

void orc_so3_exp(const double *w, double dt, double *R) {
    M3 r = Exp(V3(w[0], w[1], w[2]), dt);
    std::memcpy(R, r.m, 72);
}
void orc_so3_exp3(const double *w, double *R) {
    M3 r = Exp3(w[0], w[1], w[2]);
    std::memcpy(R, r.m, 72);
}
void orc_so3_log(const double *R, double *w) {
    M3 r;
    std::memcpy(r.m, R, 72);
    V3 l = Log(r);
    std::memcpy(w, l.v, 24);
}
void orc_jacobi6(const double *A, double *evals, double *evecs) { jacobi_eig_sym(A, 6, evals, evecs); }
int orc_inverse(const double *A, int n, double *out) {
    Mat a(A, A + (size_t)n * n), inv;
    bool ok = mat_inverse(a, n, inv);
    std::memcpy(out, inv.data(), sizeof(double) * n * n);
    return ok ? 1 : 0;
}

// Compensate() alone: state flat(36), poses: np x 22 doubles (offset, acc3, gyr3, vel3, pos3, rot9),
// pts48 sorted by normal_x on entry (the caller sorts), modified in place.
int orc_deskew_compensate(const double *state36, const double *poses, int np, float *pts48, int n) {
    State st;
    double f[36 + DIM * DIM];
    std::memset(f, 0, sizeof(f));
    std::memcpy(f, state36, 36 * sizeof(double));
    state_from_flat(st, f);
    ImuProcess ip;
    ip.IMUpose.resize(np);
    for (int i = 0; i < np; i++) {
        const double *p = poses + 22 * (size_t)i;
        Pose6D &o = ip.IMUpose[i];
        o.offset_time = p[0];
        std::memcpy(o.acc, p + 1, 24);
        std::memcpy(o.gyr, p + 4, 24);
        std::memcpy(o.vel, p + 7, 24);
        std::memcpy(o.pos, p + 10, 24);
        std::memcpy(o.rot, p + 13, 72);
    }
    std::vector<Pt> v(reinterpret_cast<Pt *>(pts48), reinterpret_cast<Pt *>(pts48) + n);
    ip.Compensate(st, v);
    std::memcpy(pts48, v.data(), (size_t)n * 48);
    return ip.first_point_repeat;
}

// ---------------------------------------------------------------- map back-ends
static MapBackend *make_map(int kind, float ds) {
    if (kind == 1) {
        if (!g_ref.ok()) return nullptr;
        return new RefMap(&g_ref, ds);
    }
    return new PortMap(ds);
}
// front end: layout7 = point_step, off_x, off_y, off_z, off_intensity, off_ring, off_time; out48: cap records
int orc_frontend_sample(const void *data, int n, const int *layout7, int sensor, int point_filter_num, float min_range, float max_range,
                        void *out48, int cap, double *timespan, double *stamp_shift) {
    orc::CloudLayout L = {layout7[0], layout7[1], layout7[2], layout7[3], layout7[4], layout7[5], layout7[6]};
    std::vector<orc::Pt> s;
    double ts = 0, sh = 0;
    int m = orc::frontend_sample(static_cast<const unsigned char *>(data), n, L, sensor, point_filter_num, min_range, max_range, s, ts, sh);
    if (timespan) *timespan = ts;
    if (stamp_shift) *stamp_shift = sh;
    for (int i = 0; i < m && i < cap; i++) std::memcpy(static_cast<char *>(out48) + 48 * (size_t)i, &s[i], 48);
    return m;
}

void *orc_map_create(int kind, float ds) { return make_map(kind, ds); }
void orc_map_destroy(void *h) { delete static_cast<MapBackend *>(h); }
void orc_map_build(void *h, const float *p4, int n) {
    std::vector<P4> v(reinterpret_cast<const P4 *>(p4), reinterpret_cast<const P4 *>(p4) + n);
    static_cast<MapBackend *>(h)->build(v);
}
void orc_map_knn(void *h, const float *q, int nq, int k, float *out_p4, float *out_d2, int *out_cnt) {
    MapBackend *m = static_cast<MapBackend *>(h);
    std::vector<P4> near;
    std::vector<float> d2;
    for (int i = 0; i < nq; i++) {
        m->knn(q[3 * i], q[3 * i + 1], q[3 * i + 2], k, near, d2);
        out_cnt[i] = (int)near.size();
        for (int j = 0; j < k; j++) {
            float *o = out_p4 + ((size_t)i * k + j) * 4;
            if (j < (int)near.size()) {
                o[0] = near[j].x;
                o[1] = near[j].y;
                o[2] = near[j].z;
                o[3] = near[j].w;
                out_d2[(size_t)i * k + j] = d2[j];
            } else {
                o[0] = o[1] = o[2] = o[3] = 0.f;
                out_d2[(size_t)i * k + j] = -1.f;
            }
        }
    }
}
int orc_map_add(void *h, const float *p4, int n, int ds_on) {
    std::vector<P4> v(reinterpret_cast<const P4 *>(p4), reinterpret_cast<const P4 *>(p4) + n);
    return static_cast<MapBackend *>(h)->add_points(v, ds_on != 0);
}
int orc_map_delete_boxes(void *h, const float *boxes, int nb) {
    std::vector<Box> b(nb);
    for (int i = 0; i < nb; i++) {
        std::memcpy(b[i].mn, boxes + 6 * i, 12);
        std::memcpy(b[i].mx, boxes + 6 * i + 3, 12);
    }
    return static_cast<MapBackend *>(h)->delete_boxes(b);
}
int orc_map_validnum(void *h) { return static_cast<MapBackend *>(h)->validnum(); }
int orc_map_flatten(void *h, float *out, int cap) {
    std::vector<P4> v;
    static_cast<MapBackend *>(h)->flatten(v);
    int n = (int)v.size();
    if (out)
        for (int i = 0; i < n && i < cap; i++) std::memcpy(out + 4 * (size_t)i, &v[i], 16);
    return n;
}

// ---------------------------------------------------------------- the per-scan pipeline
struct OrcLioConfig {
    int max_iteration;
    double filter_size_surf, filter_size_map, cube_len;
    int extrinsic_est_en;
    int featptsThreshold;
    double beta;
    float det_range;
    double extrinT[3];
    double extrinR[9];
};
struct OrcThermal {
    int tis_online, recv_n;
    double delta_pos[3], delta_quat[4], delta_vel[3], cov_slots[8];
    double l2l_pos[3], l2l_quat[4], l2l_vel[3], l2l_cov_slots[8];
};
struct OrcScanSummary {
    int had_points, built_map, did_update, ekf_stop;
    int n_raw, n_down, map_points_before, deleted, added, n_iters;
    int n_added_ds, n_added_raw;
    double eigvals[6], eigvecs[36];
    double state_prop[36];
    double t_deskew, t_voxel, t_knn, t_resid, t_solve, t_insert, t_delete;
};
struct OrcIter {
    int iter, effct_feat_num, converged, ekf_stop, did_match, n_down;
    double total_residual, res_mean_last;
    double HtH[144], Htr[12], pose_in[24], state_out[36], solution[24];
};

struct LioHandle {
    Lio *lio;
    ScanResult last;
};

void *orc_lio_create(const OrcLioConfig *c, int map_kind) {
    LioConfig cfg;
    cfg.max_iteration = c->max_iteration;
    cfg.filter_size_surf = c->filter_size_surf;
    cfg.filter_size_map = c->filter_size_map;
    cfg.cube_len = c->cube_len;
    cfg.extrinsic_est_en = c->extrinsic_est_en != 0;
    cfg.featptsThreshold = c->featptsThreshold;
    cfg.beta = c->beta;
    cfg.det_range = c->det_range;
    cfg.extrinT = V3(c->extrinT[0], c->extrinT[1], c->extrinT[2]);
    std::memcpy(cfg.extrinR.m, c->extrinR, 72);
    MapBackend *m = make_map(map_kind, (float)cfg.filter_size_map);  // set_downsample_param, laserMapping.cpp:784
    if (!m) return nullptr;
    LioHandle *h = new LioHandle;
    h->lio = new Lio(cfg, m);
    return h;
}
void orc_lio_destroy(void *hh) {
    LioHandle *h = static_cast<LioHandle *>(hh);
    delete h->lio;
    delete h;
}
void orc_lio_set_threads(void *hh, int n) { static_cast<LioHandle *>(hh)->lio->omp_threads = n < 1 ? 1 : n; }
void orc_lio_on_lidar_msg(void *hh) { static_cast<LioHandle *>(hh)->lio->on_lidar_msg(); }
void orc_lio_on_edge_count(void *hh, int n) { static_cast<LioHandle *>(hh)->lio->on_edge_count(n); }
// skip the ~100-sample IMU initialisation (IMU_Processing.hpp:385-405) with a given mean_acc
void orc_lio_force_imu_ready(void *hh, const double *mean_acc, const double *last_imu7) {
    Lio *l = static_cast<LioHandle *>(hh)->lio;
    l->imu.imu_need_init_ = false;
    l->imu.b_first_frame_ = false;
    l->imu.init_iter_num = MAX_INI_COUNT + 1;
    l->imu.mean_acc = V3(mean_acc[0], mean_acc[1], mean_acc[2]);
    l->imu.last_imu_.t = last_imu7[0];
    std::memcpy(l->imu.last_imu_.acc, last_imu7 + 1, 24);
    std::memcpy(l->imu.last_imu_.gyr, last_imu7 + 4, 24);
}
void orc_lio_get_state(void *hh, double *out612) { state_to_flat(static_cast<LioHandle *>(hh)->lio->state, out612); }
void orc_lio_set_state(void *hh, const double *in612) {
    Lio *l = static_cast<LioHandle *>(hh)->lio;
    state_from_flat(l->state, in612);
}
void orc_lio_set_last_state(void *hh, const double *in612) {
    Lio *l = static_cast<LioHandle *>(hh)->lio;
    state_from_flat(l->last_state, in612);
    state_from_flat(l->last_nodegared_state, in612);
}
void *orc_lio_map(void *hh) { return static_cast<LioHandle *>(hh)->lio->map.get(); }
int orc_lio_imu_ready(void *hh) { return static_cast<LioHandle *>(hh)->lio->imu.imu_need_init_ ? 0 : 1; }

int orc_lio_process_scan(void *hh, const float *pts48, int n, double lidar_beg_time, const double *imu7, int m,
                         const OrcThermal *th, OrcScanSummary *out) {
    LioHandle *h = static_cast<LioHandle *>(hh);
    MeasureGroup meas;
    meas.lidar_beg_time = lidar_beg_time;
    meas.lidar.assign(reinterpret_cast<const Pt *>(pts48), reinterpret_cast<const Pt *>(pts48) + n);
    meas.observation_end_time = lidar_beg_time + (n > 0 ? (double)meas.lidar.back().nz : 0.0);  // laserMapping.cpp:546
    meas.imu.resize(m);
    for (int i = 0; i < m; i++) {
        meas.imu[i].t = imu7[7 * i];
        std::memcpy(meas.imu[i].acc, imu7 + 7 * i + 1, 24);
        std::memcpy(meas.imu[i].gyr, imu7 + 7 * i + 4, 24);
    }
    ThermalInputs t;
    if (th) {
        t.tis_online = th->tis_online != 0;
        t.recv_n = th->recv_n;
        std::memcpy(t.delta_pos, th->delta_pos, sizeof(t.delta_pos));
        std::memcpy(t.delta_quat, th->delta_quat, sizeof(t.delta_quat));
        std::memcpy(t.delta_vel, th->delta_vel, sizeof(t.delta_vel));
        std::memcpy(t.cov_slots, th->cov_slots, sizeof(t.cov_slots));
        std::memcpy(t.l2l_pos, th->l2l_pos, sizeof(t.l2l_pos));
        std::memcpy(t.l2l_quat, th->l2l_quat, sizeof(t.l2l_quat));
        std::memcpy(t.l2l_vel, th->l2l_vel, sizeof(t.l2l_vel));
        std::memcpy(t.l2l_cov_slots, th->l2l_cov_slots, sizeof(t.l2l_cov_slots));
    }
    h->lio->process_scan(meas, t, h->last);
    const ScanResult &r = h->last;
    std::memset(out, 0, sizeof(*out));
    out->had_points = r.had_points;
    out->built_map = r.built_map;
    out->did_update = r.did_update;
    out->ekf_stop = r.ekf_stop;
    out->n_raw = r.n_raw;
    out->n_down = r.n_down;
    out->map_points_before = r.map_points_before;
    out->deleted = r.deleted;
    out->added = r.added;
    out->n_iters = (int)r.iters.size();
    out->n_added_ds = (int)r.added_ds.size();
    out->n_added_raw = (int)r.added_raw.size();
    std::memcpy(out->eigvals, r.eigvals, sizeof(out->eigvals));
    std::memcpy(out->eigvecs, r.eigvecs, sizeof(out->eigvecs));
    std::memcpy(out->state_prop, r.state_prop, sizeof(out->state_prop));
    out->t_deskew = h->lio->t_deskew;
    out->t_voxel = h->lio->t_voxel;
    out->t_knn = h->lio->t_knn;
    out->t_resid = h->lio->t_resid;
    out->t_solve = h->lio->t_solve;
    out->t_insert = h->lio->t_insert;
    out->t_delete = h->lio->t_delete;
    return 0;
}
int orc_lio_get_iters(void *hh, OrcIter *out, int cap) {
    const ScanResult &r = static_cast<LioHandle *>(hh)->last;
    int n = (int)r.iters.size();
    for (int i = 0; i < n && i < cap; i++) {
        const IterRecord &s = r.iters[i];
        OrcIter &o = out[i];
        o.iter = s.iter;
        o.effct_feat_num = s.effct_feat_num;
        o.converged = s.converged;
        o.ekf_stop = s.ekf_stop;
        o.did_match = s.did_match;
        o.n_down = s.n_down;
        o.total_residual = s.total_residual;
        o.res_mean_last = s.res_mean_last;
        std::memcpy(o.HtH, s.HtH, sizeof(o.HtH));
        std::memcpy(o.Htr, s.Htr, sizeof(o.Htr));
        std::memcpy(o.pose_in, s.pose_in, sizeof(o.pose_in));
        std::memcpy(o.state_out, s.state_out, sizeof(o.state_out));
        std::memcpy(o.solution, s.solution, sizeof(o.solution));
    }
    return n;
}
int orc_lio_get_undistort(void *hh, float *out48, int cap) {
    const ScanResult &r = static_cast<LioHandle *>(hh)->last;
    int n = (int)r.feats_undistort.size();
    if (out48) std::memcpy(out48, r.feats_undistort.data(), (size_t)std::min(n, cap) * 48);
    return n;
}
int orc_lio_get_feats_down(void *hh, float *out48, int cap) {
    const ScanResult &r = static_cast<LioHandle *>(hh)->last;
    int n = (int)r.feats_down.size();
    if (out48) std::memcpy(out48, r.feats_down.data(), (size_t)std::min(n, cap) * 48);
    return n;
}
// nearest: n_down x 5 x 4 floats, d2: n_down x 5, cnt: n_down, selected: n_down
int orc_lio_get_nearest(void *hh, float *near, float *d2, int *cnt, uint8_t *selected, int cap) {
    const ScanResult &r = static_cast<LioHandle *>(hh)->last;
    int n = (int)r.nearest_cnt.size();
    int m = std::min(n, cap);
    if (m > 0) {
        std::memcpy(near, r.nearest.data(), (size_t)m * 5 * 16);
        std::memcpy(d2, r.nearest_d2.data(), (size_t)m * 5 * 4);
        std::memcpy(cnt, r.nearest_cnt.data(), (size_t)m * 4);
        std::memcpy(selected, r.selected.data(), (size_t)m);
    }
    return n;
}
int orc_lio_get_added(void *hh, float *ds_p4, int cap_ds, float *raw_p4, int cap_raw) {
    const ScanResult &r = static_cast<LioHandle *>(hh)->last;
    if (ds_p4) std::memcpy(ds_p4, r.added_ds.data(), (size_t)std::min((int)r.added_ds.size(), cap_ds) * 16);
    if (raw_p4) std::memcpy(raw_p4, r.added_raw.data(), (size_t)std::min((int)r.added_raw.size(), cap_raw) * 16);
    return (int)(r.added_ds.size() + r.added_raw.size());
}
// IMUpose of the last Process call: np x 22 doubles
int orc_lio_get_imu_poses(void *hh, double *out, int cap) {
    const Lio *l = static_cast<LioHandle *>(hh)->lio;
    int n = (int)l->imu.IMUpose.size();
    for (int i = 0; i < n && i < cap; i++) {
        const Pose6D &p = l->imu.IMUpose[i];
        double *o = out + 22 * (size_t)i;
        o[0] = p.offset_time;
        std::memcpy(o + 1, p.acc, 24);
        std::memcpy(o + 4, p.gyr, 24);
        std::memcpy(o + 7, p.vel, 24);
        std::memcpy(o + 10, p.pos, 24);
        std::memcpy(o + 13, p.rot, 72);
    }
    return n;
}
void orc_lio_get_flags(void *hh, int *out8) {
    const Lio *l = static_cast<LioHandle *>(hh)->lio;
    out8[0] = l->EKF_stop_flg;
    out8[1] = l->flg_EKF_inited;
    out8[2] = l->dynamic_effect_featurepoints_threshold;
    out8[3] = (int)l->lidar_cnt;
    out8[4] = l->Localmap_Initialized;
    out8[5] = (int)l->effct_q.size();
    out8[6] = l->imu.first_point_repeat;
    out8[7] = 0;
}
void orc_lio_get_localmap(void *hh, float *out6) {
    const Lio *l = static_cast<LioHandle *>(hh)->lio;
    std::memcpy(out6, l->LocalMap_Points.mn, 12);
    std::memcpy(out6 + 3, l->LocalMap_Points.mx, 12);
}

}  // extern "C"
