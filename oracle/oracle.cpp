// oracle/oracle.cpp -- TEST INFRASTRUCTURE ONLY (CPU parity oracle, kind "port").
//
// Restatement of the per-scan body of eskf_lio's main loop
// (/root/reference/eskf_lio/src/laserMapping.cpp:731-1177) on top of the pieces in
// oracle.hpp.  The explicit E x 12 Jacobian `Hsub`, the explicit gain K = K_1[:, :12] H^T
// and `solution = K z + vec - K H vec_12` are kept in the reference's literal form so
// that the product's reduced form (H^T H, H^T r only) is checked against something that
// was not derived the same way.
#include "oracle.hpp"

#include <chrono>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

double now_sec() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void Lio::process_scan(const MeasureGroup &meas, const ThermalInputs &th, ScanResult &res) {
    const int NUM_MATCH_POINTS = 5;
    const double LASER_POINT_COV = 0.0015;  // laserMapping.cpp:76
    const int QUEUE_SIZE = 10;              // laserMapping.cpp:192
    const int NUM_MAX_ITERATIONS = cfg.max_iteration;
    res = ScanResult();
    t_deskew = t_voxel = t_knn = t_resid = t_solve = t_insert = t_delete = 0;

    if (flg_first_scan) {  // :736-740
        first_lidar_time = meas.lidar_beg_time;
        flg_first_scan = false;
    }
    double t0 = now_sec();
    std::vector<Pt> feats_undistort;
    imu.Process(meas, state, feats_undistort, EKF_stop_flg);  // :750
    t_deskew = now_sec() - t0;
    State state_propagat = state;                                     // :752
    V3 pos_lid = state_propagat.pos_end + state_propagat.rot_end * state_propagat.T_L_I;  // :753
    {
        double f[36 + DIM * DIM];
        state_to_flat(state_propagat, f);
        std::memcpy(res.state_prop, f, sizeof(res.state_prop));
    }
    res.n_raw = (int)feats_undistort.size();
    if (feats_undistort.empty()) return;  // :755-759
    res.had_points = true;
    flg_EKF_inited = true;  // :762 with INIT_TIME == 0 (:75)

    t0 = now_sec();
    std::vector<Box> cub_needrm;
    res.deleted = fov_segment(pos_lid, cub_needrm);  // :772
    t_delete = now_sec() - t0;

    t0 = now_sec();
    std::vector<Pt> feats_down;
    voxel_grid(feats_undistort.data(), (int)feats_undistort.size(), (float)cfg.filter_size_surf, feats_down, false);  // :775-776
    t_voxel = now_sec() - t0;
    int feats_down_size = (int)feats_down.size();
    res.n_down = feats_down_size;
    res.feats_undistort = feats_undistort;
    res.feats_down = feats_down;

    if (!map->has_root()) {  // :780-793
        if (feats_down_size > 5) {
            std::vector<P4> w(feats_down_size);
            for (int i = 0; i < feats_down_size; i++) {
                float o[3];
                body_to_world(state, feats_down[i].x, feats_down[i].y, feats_down[i].z, o);
                w[i] = {o[0], o[1], o[2], feats_down[i].intensity};
            }
            map->build(w);
            res.built_map = true;
        }
        return;
    }
    int featsFromMapNum = map->validnum();  // :794
    res.map_points_before = featsFromMapNum;

    std::vector<float> coeff(4 * (size_t)feats_down_size, 0.f);  // coeffSel_tmpt x,y,z,intensity
    std::vector<double> res_last(feats_down_size, 1000.0);        // :802
    int effct_feat_num = 0;

    if (featsFromMapNum >= 5) {  // :804
        res.did_update = true;
        Nearest_Points.assign(feats_down_size, std::vector<P4>());
        std::vector<std::vector<float>> Nearest_D2(feats_down_size);
        std::vector<uint8_t> point_selected_surf(feats_down_size, 1);  // :812
        std::vector<float> world(3 * (size_t)feats_down_size);
        int rematch_num = 0;
        bool rematch_en = false;
        bool flg_EKF_converged = false;
        Mat Hsub, K;
        std::vector<double> meas_vec;
        std::vector<int> eff_idx;

        for (int iterCount = 0; iterCount < NUM_MAX_ITERATIONS; iterCount++) {  // :820
            IterRecord rec;
            rec.iter = iterCount;
            rec.n_down = feats_down_size;
            rec.did_match = (iterCount == 0 || rematch_en);
            std::memcpy(rec.pose_in, state.rot_end.m, 72);
            std::memcpy(rec.pose_in + 9, state.pos_end.v, 24);
            std::memcpy(rec.pose_in + 12, state.R_L_I.m, 72);
            std::memcpy(rec.pose_in + 21, state.T_L_I.v, 24);

            /** closest surface search and residual computation **/  // :826-882
            double tk = 0, tr = 0;
            double tA = now_sec();
            if (rec.did_match) {
#pragma omp parallel for num_threads(omp_threads) schedule(static) if (omp_threads > 1)
                for (int i = 0; i < feats_down_size; i++) {
                    float w[3];
                    body_to_world(state, feats_down[i].x, feats_down[i].y, feats_down[i].z, w);
                    map->knn(w[0], w[1], w[2], NUM_MATCH_POINTS, Nearest_Points[i], Nearest_D2[i]);
                    point_selected_surf[i] = (int)Nearest_Points[i].size() < NUM_MATCH_POINTS ? 0
                                             : Nearest_D2[i][NUM_MATCH_POINTS - 1] > 5   ? 0
                                                                                         : 1;
                }
            }
            tk = now_sec() - tA;
            tA = now_sec();
#pragma omp parallel for num_threads(omp_threads) schedule(static) if (omp_threads > 1)
            for (int i = 0; i < feats_down_size; i++) {
                const Pt &pb = feats_down[i];
                V3 p_body(pb.x, pb.y, pb.z);
                V3 p_global = state.rot_end * (state.R_L_I * p_body + state.T_L_I) + state.pos_end;
                float wx = (float)p_global[0], wy = (float)p_global[1], wz = (float)p_global[2];
                world[3 * i] = wx;
                world[3 * i + 1] = wy;
                world[3 * i + 2] = wz;
                if (!point_selected_surf[i]) continue;  // :857
                float pabcd[4];
                point_selected_surf[i] = 0;
                const std::vector<P4> &pn = Nearest_Points[i];
                float pts[5][3];
                for (int j = 0; j < 5; j++) {
                    pts[j][0] = pn[j].x;
                    pts[j][1] = pn[j].y;
                    pts[j][2] = pn[j].z;
                }
                if (esti_plane(pabcd, pts, 0.1f)) {  // :863
                    float pd2 = pabcd[0] * wx + pabcd[1] * wy;  // :866
                    pd2 = pd2 + pabcd[2] * wz;
                    pd2 = pd2 + pabcd[3];
                    float s = (float)(1 - 0.9 * (double)std::fabs(pd2) / std::sqrt(norm(p_body)));  // :868
                    if ((double)s > 0.9) {
                        point_selected_surf[i] = 1;
                        coeff[4 * i] = pabcd[0];
                        coeff[4 * i + 1] = pabcd[1];
                        coeff[4 * i + 2] = pabcd[2];
                        coeff[4 * i + 3] = pd2;
                        res_last[i] = (double)std::fabs(pd2);
                    }
                }
            }
            double total_residual = 0.0;  // :884-896
            effct_feat_num = 0;
            eff_idx.clear();
            for (int i = 0; i < feats_down_size; i++) {
                if (point_selected_surf[i] && (res_last[i] <= 2.0)) {
                    eff_idx.push_back(i);
                    total_residual += res_last[i];
                    effct_feat_num++;
                }
            }
            effct_q.push_back(effct_feat_num);  // :899-918
            if ((int)effct_q.size() > QUEUE_SIZE) effct_q.pop_front();
            EKF_stop_flg = false;
            for (int v : effct_q)
                if (v <= dynamic_effect_featurepoints_threshold) {
                    EKF_stop_flg = true;
                    break;
                }
            double res_mean_last = total_residual / effct_feat_num;  // :932 (inf/nan when 0, as in the reference)

            /*** Measurement Jacobian H and measurement vector ***/  // :942-979
            Hsub.assign((size_t)effct_feat_num * 12, 0.0);
            meas_vec.assign(effct_feat_num, 0.0);
            M3 rotT = transpose(state.rot_end), RLIT = transpose(state.R_L_I);
            for (int e = 0; e < effct_feat_num; e++) {
                int i = eff_idx[e];
                V3 point_this_be(feats_down[i].x, feats_down[i].y, feats_down[i].z);
                M3 point_be_crossmat = skew(point_this_be);
                V3 point_this = state.R_L_I * point_this_be + state.T_L_I;
                M3 point_crossmat = skew(point_this);
                V3 norm_vec(coeff[4 * i], coeff[4 * i + 1], coeff[4 * i + 2]);
                V3 C = rotT * norm_vec;
                V3 A = point_crossmat * C;
                double *row = &Hsub[(size_t)e * 12];
                row[0] = A[0];
                row[1] = A[1];
                row[2] = A[2];
                row[3] = coeff[4 * i];
                row[4] = coeff[4 * i + 1];
                row[5] = coeff[4 * i + 2];
                if (cfg.extrinsic_est_en) {
                    V3 B = (point_be_crossmat * RLIT) * C;
                    row[6] = B[0];
                    row[7] = B[1];
                    row[8] = B[2];
                    row[9] = C[0];
                    row[10] = C[1];
                    row[11] = C[2];
                }
                meas_vec[e] = -(double)coeff[4 * i + 3];
            }
            tr = now_sec() - tA;
            t_knn += tk;
            t_resid += tr;

            // reduced forms recorded for the parity tests
            tA = now_sec();
            Mat HsubT = mat_T(Hsub, effct_feat_num, 12);
            Mat HtH12 = mat_mul(HsubT, 12, effct_feat_num, Hsub, 12);
            std::memcpy(rec.HtH, HtH12.data(), sizeof(rec.HtH));
            for (int a = 0; a < 12; a++) {
                double s = 0;
                for (int e = 0; e < effct_feat_num; e++) s += Hsub[(size_t)e * 12 + a] * meas_vec[e];
                rec.Htr[a] = s;
            }
            rec.effct_feat_num = effct_feat_num;
            rec.total_residual = total_residual;
            rec.res_mean_last = res_mean_last;
            rec.ekf_stop = EKF_stop_flg;
            std::memset(rec.solution, 0, sizeof(rec.solution));

            /*** Iterative Kalman Filter Update ***/  // :985-1063
            // (the `!flg_EKF_inited && !EKF_stop_flg` initialisation branch at :986-1011 is
            //  unreachable: INIT_TIME is 0 (:75,:762) and the only writer of
            //  flg_EKF_inited=false (:1062) is followed by a break (:1095-1100).)
            if (!EKF_stop_flg) {
                Mat H_T_H((size_t)DIM * DIM, 0.0);
                for (int a = 0; a < 12; a++)
                    for (int b = 0; b < 12; b++) H_T_H[(size_t)a * DIM + b] = HtH12[(size_t)a * 12 + b];
                Mat Pn((size_t)DIM * DIM);
                for (int i = 0; i < DIM * DIM; i++) Pn[i] = state.cov[i] / LASER_POINT_COV;
                Mat Pinv, K1;
                mat_inverse(Pn, DIM, Pinv);
                for (int i = 0; i < DIM * DIM; i++) Pinv[i] += H_T_H[i];
                mat_inverse(Pinv, DIM, K1);  // :1017-1018
                Mat K1c((size_t)DIM * 12);
                for (int i = 0; i < DIM; i++)
                    for (int j = 0; j < 12; j++) K1c[(size_t)i * 12 + j] = K1[(size_t)i * DIM + j];
                K = mat_mul(K1c, DIM, 12, HsubT, effct_feat_num);  // :1019, 24 x E
                double vec[DIM];
                state_minus(state_propagat, state, vec);  // :1028
                // solution = K * meas_vec + vec - K * Hsub * vec.block<12,1>(0,0)   :1032
                Mat KH = mat_mul(K, DIM, effct_feat_num, Hsub, 12);  // 24 x 12
                double solution[DIM];
                for (int i = 0; i < DIM; i++) {
                    double kz = 0;
                    for (int e = 0; e < effct_feat_num; e++) kz += K[(size_t)i * effct_feat_num + e] * meas_vec[e];
                    double khv = 0;
                    for (int j = 0; j < 12; j++) khv += KH[(size_t)i * 12 + j] * vec[j];
                    solution[i] = kz + vec[i] - khv;
                }
                state_add_inplace(state, solution);  // :1033
                std::memcpy(rec.solution, solution, sizeof(solution));
                double rn = std::sqrt(solution[0] * solution[0] + solution[1] * solution[1] + solution[2] * solution[2]);
                double tn = std::sqrt(solution[3] * solution[3] + solution[4] * solution[4] + solution[5] * solution[5]);
                flg_EKF_converged = false;
                if ((rn * 57.3 < 0.01) && (tn * 100 < 0.015)) flg_EKF_converged = true;  // :1040
                last_nodegared_state = state;  // :1050
            } else {
                // :1054-1063
                State d = odomToState(th.delta_pos, th.delta_quat, th.delta_vel, th.cov_slots);
                state = state_plus_state(last_nodegared_state, d);
                flg_EKF_inited = false;
            }
            rec.converged = flg_EKF_converged;
            {
                double f[36 + DIM * DIM];
                state_to_flat(state, f);
                std::memcpy(rec.state_out, f, sizeof(rec.state_out));
            }
            t_solve += now_sec() - tA;
            res.iters.push_back(rec);

            /*** Rematch Judgement ***/  // :1069-1076
            rematch_en = false;
            if (flg_EKF_converged || ((rematch_num == 0) && (iterCount == (NUM_MAX_ITERATIONS - 2)))) {
                rematch_en = true;
                rematch_num++;
            }
            /*** Convergence Judgements and Covariance Update ***/  // :1078-1101
            if (rematch_num >= 2 || (iterCount == NUM_MAX_ITERATIONS - 1)) {
                if (flg_EKF_inited) {
                    // G.block<24,12>(0,0) = K * Hsub; cov = (I - G) * cov    :1084-1085
                    Mat KH = mat_mul(K, DIM, effct_feat_num, Hsub, 12);
                    Mat ImG((size_t)DIM * DIM, 0.0);
                    for (int i = 0; i < DIM; i++) ImG[(size_t)i * DIM + i] = 1.0;
                    for (int i = 0; i < DIM; i++)
                        for (int j = 0; j < 12; j++) ImG[(size_t)i * DIM + j] -= KH[(size_t)i * 12 + j];
                    Mat P(state.cov, state.cov + DIM * DIM);
                    Mat Pn = mat_mul(ImG, DIM, DIM, P, DIM);
                    std::memcpy(state.cov, Pn.data(), sizeof(state.cov));
                }
                break;
            } else if (EKF_stop_flg) {
                break;
            }
        }

        // NEW output: eigen-decomposition of the 6x6 pose block of the last H^T H
        if (!res.iters.empty()) {
            double A6[36];
            for (int a = 0; a < 6; a++)
                for (int b = 0; b < 6; b++) A6[a * 6 + b] = res.iters.back().HtH[a * 12 + b];
            jacobi_eig_sym(A6, 6, res.eigvals, res.eigvecs);
        }

        /*** zeta blend ***/  // :1105-1129
        if ((lidar_cnt < 100) || (!th.tis_online) || (th.tis_online && lidar_cnt % 2 == 1)) {
            double alpha_l = effct_feat_num / (cfg.beta * 65536.0);
            zeta_l = 2.0 / (1.0 + std::exp(-alpha_l)) - 1;
            double zeta_l_norm = zeta_l / (zeta_l + zeta_t);
            if (lidar_cnt < 100) zeta_l_norm = 1;
            double v1[DIM], v2[DIM];
            state_minus(state_propagat, last_state, v1);
            state_minus(state, last_state, v2);
            for (int i = 0; i < DIM; i++) {
                v1[i] *= (1 - zeta_l_norm);
                v2[i] *= zeta_l_norm;
            }
            state = state_plus(state_plus(last_state, v1), v2);  // :1119 (cov := last_state.cov)
        } else {
            double zeta_t_norm = zeta_t / (zeta_l + zeta_t);
            double v1[DIM];
            state_minus(state, last_state, v1);
            for (int i = 0; i < DIM; i++) v1[i] *= (1 - zeta_t_norm);
            State o = odomToState(th.l2l_pos, th.l2l_quat, th.l2l_vel, th.l2l_cov_slots);
            state = state_plus_state(state_plus(last_state, v1), state_scale(o, zeta_t_norm));  // :1127
        }
        last_state = state;  // :1131

        // snapshot of the last match pass (what map_incremental consumes)
        res.nearest.assign((size_t)feats_down_size * 5, P4{0, 0, 0, 0});
        res.nearest_d2.assign((size_t)feats_down_size * 5, -1.f);
        res.nearest_cnt.assign(feats_down_size, 0);
        for (int i = 0; i < feats_down_size; i++) {
            res.nearest_cnt[i] = (int)Nearest_Points[i].size();
            for (int j = 0; j < (int)Nearest_Points[i].size() && j < 5; j++) {
                res.nearest[(size_t)i * 5 + j] = Nearest_Points[i][j];
                res.nearest_d2[(size_t)i * 5 + j] = Nearest_D2[i][j];
            }
        }
        res.selected = point_selected_surf;

        /*** add new frame points to map ***/  // :1164-1168
        double tA = now_sec();
        if (!EKF_stop_flg) map_incremental(feats_down, res);
        t_insert = now_sec() - tA;
    }
    res.ekf_stop = EKF_stop_flg;
}

}  // namespace orc
