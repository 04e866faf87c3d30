"""Parity checks at benchmark scale -- TEST INFRASTRUCTURE (used by tests/ and by bench.py's cpu_baseline leg only).

Two layers, both against the oracle (reference ikd-Tree compiled from /root/reference + the restated loop):

  identical inputs   the oracle's own feats_down, per-iteration pose and map-before are fed to the device measurement
                     model (dlt_measure): neighbour sets and point_selected_surf bit-exact, effct_feat_num exact,
                     H^T H / H^T r relative error, map_incremental add lists and map contents equal as sets.
  re-synced pipeline before every scan the device pipeline (dlt_lio_process_scan: deskew -> VoxelGrid -> loop -> blend ->
                     map_incremental) is given the oracle's state and holds the oracle's map; the pose after ONE scan from
                     identical raw inputs is compared (north star: 1e-5 relative), together with how many VoxelGrid voxels
                     differ (CUDA vs glibc sin/cos in the deskew move a point across a voxel face now and then).

`chained=True` drops the re-sync so the two chains drift on their own maps: that is the series the chained-replay
tolerances in tests/test_pipeline_parity.py are derived from (tools/parity_series.py).
"""
from __future__ import annotations

import numpy as np


def sort_rows(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(a, np.float32)[:, :3])
    if len(a) == 0:
        return a
    v = a.view(np.uint32).astype(np.uint64)
    # lexicographic order over the bit patterns: any total order will do for a set comparison
    order = np.lexsort((v[:, 2], v[:, 1], v[:, 0]))
    return a[order]


def map_set_diff(a: np.ndarray, b: np.ndarray) -> int:
    """size of the symmetric difference of two point sets (xyz compared bit for bit; multiset semantics)"""
    sa, sb = sort_rows(a), sort_rows(b)
    if sa.shape == sb.shape and np.array_equal(sa.view(np.uint32), sb.view(np.uint32)):
        return 0
    ka = {}
    for row in map(bytes, sa.view(np.uint8).reshape(len(sa), 12)):
        ka[row] = ka.get(row, 0) + 1
    diff = 0
    for row in map(bytes, sb.view(np.uint8).reshape(len(sb), 12)):
        c = ka.get(row, 0)
        if c:
            ka[row] = c - 1
        else:
            diff += 1
    return diff + sum(ka.values())


def map_quantised_diff(a: np.ndarray, b: np.ndarray, quantum: float = 1e-3) -> int:
    """points of one set with no partner within `quantum` metres in the other, after the bit-identical ones are paired off:
    the pipeline's inserted points are centroids whose last bit may differ from the oracle's (fixed-point vs float sums)"""
    from scipy.spatial import cKDTree

    def rows(p):
        q = np.ascontiguousarray(np.asarray(p, np.float32)[:, :3])
        return q, q.view(np.dtype((np.void, 12))).ravel()
    pa, va = rows(a)
    pb, vb = rows(b)
    ra = pa[~np.isin(va, vb)]
    rb = pb[~np.isin(vb, va)]
    if len(ra) == 0 or len(rb) == 0:
        return int(len(ra) + len(rb))
    d, idx = cKDTree(rb.astype(np.float64)).query(ra.astype(np.float64), k=1, distance_upper_bound=quantum)
    hit = np.isfinite(d)
    return int((~hit).sum() + len(rb) - len(np.unique(idx[hit])))


def voxel_keys(xyz: np.ndarray, leaf: float = 0.5) -> np.ndarray:
    """absolute integer voxel coordinates of VoxelGrid centroids (a centroid lies inside its voxel), packed"""
    ijk = np.floor(np.asarray(xyz, np.float32)[:, :3] * np.float32(1.0 / leaf)).astype(np.int64)
    return (ijk[:, 0] + (1 << 20)) | ((ijk[:, 1] + (1 << 20)) << 21) | ((ijk[:, 2] + (1 << 20)) << 42)


def pose_rel_err(a, b) -> float:
    """max(|dR| entries, |dp| / max(1, |p|)): the north star's relative pose error"""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    rot = float(np.abs(a[0:9] - b[0:9]).max())
    pos = float(np.abs(a[9:12] - b[9:12]).max() / max(1.0, np.abs(b[9:12]).max()))
    return max(rot, pos)


def rel_err(a, b) -> float:
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


class ParityRun:
    """Drives an oracle pipeline and the device side by side over a scan sequence and accumulates the parity record."""

    def __init__(self, lib, oracle, seq, map_pts, lm_kwargs=None, max_scan_points=1 << 18, threads=1, chained=False, identical=True,
                 device=0, device_loop=-1, check_maps=True, featptsThreshold=30):
        import helpers
        import oracle_binding as ob
        from daliti_b200 import synth
        from daliti_b200.binding import ScanToMap
        from daliti_b200.lio import LaserMapping

        self.lib, self.oracle, self.seq = lib, oracle, seq
        self.chained, self.identical, self.check_maps = chained, identical, check_maps
        lm_kwargs = dict(lm_kwargs or {})
        kind = ob.MAP_REF if oracle.ref_ok else ob.MAP_PORT
        self.kind = "reference" if oracle.ref_ok else "port"
        ocfg = {k: v for k, v in lm_kwargs.items() if k in ("cube_len", "det_range", "max_iteration")}
        self.lio = oracle.new_lio(ob.default_lio_config(featptsThreshold=featptsThreshold, **ocfg), kind)
        mean_acc = [0.0, 0.0, synth.G]
        last_imu = np.concatenate([[seq.t_start - 0.005], [0, 0, synth.G], [0, 0, seq.traj.yaw_rate]])
        s0 = helpers.state612(seq.traj, seq.t_start)
        self.lio.force_imu_ready(mean_acc, last_imu)
        self.lio.set_state(s0)
        self.lio.set_threads(threads)
        n_map = len(map_pts) if map_pts is not None else 0
        cap = max(1 << 18, 2 * n_map + (1 << 16))
        self.lm = LaserMapping(lib, dev=dict(device=device, max_scan_points=max_scan_points, max_map_points=cap), featptsThreshold=featptsThreshold,
                               device_loop=device_loop, **lm_kwargs)
        self.lm.force_imu_ready(mean_acc, last_imu)
        self.lm.set_state(s0)
        self.dm = None
        if map_pts is not None:
            self.lio.map().build(map_pts)
            self.lm.device.map_build(map_pts)
            if identical:
                self.dm = ScanToMap(lib, device=device, max_scan_points=max_scan_points, max_map_points=cap)
                self.dm.map_build(map_pts)
        self.rec = dict(scans=0, knn_queries=0, knn_sets_equal=True, selected_equal=True, effct_equal=True, add_lists_equal=True,
                        map_contents_equal=True, max_HtH_rel_err=0.0, max_Htr_rel_err=0.0, max_pose_rel_err=0.0, max_iter_pose_rel_err=0.0,
                        n_iters_equal=True, voxels_differing=0, voxels_total=0, pipeline_effct_max_abs_diff=0, pipeline_added_max_abs_diff=0,
                        pipeline_map_points_differing_max=0, pipeline_map_diff_quantum_m=1e-3, resyncs=0, oracle_kind=self.kind, mode="chained" if chained else "re-synced per scan")
        self.series = []
        self.cpu_times = []

    def close(self):
        self.lm.close()
        if self.dm is not None:
            self.dm.close()
        self.lio.close()

    def step(self, k, scan=None):
        import time

        lio, lm, dm, rec = self.lio, self.lm, self.dm, self.rec
        pts, t_beg, imu = scan if scan is not None else self.seq.scan(k)
        if not self.chained:  # identical state on both sides (state, last_state, last_nodegared_state)
            s = lio.get_state()
            lio.set_state(s, also_last=True)
            lm.set_state(s, also_last=True)
        lio.on_lidar_msg()
        lm.on_lidar_msg()
        t0 = time.perf_counter()
        so = lio.process_scan(pts, t_beg, imu)
        self.cpu_times.append((time.perf_counter() - t0, so.n_raw))
        sd = lm.process_scan(pts, t_beg, imu)
        row = dict(scan=k, n_down_dev=sd.n_down, n_down_orc=so.n_down, n_iters_dev=sd.n_iters, n_iters_orc=so.n_iters)
        rec["scans"] += 1
        if (sd.had_points, sd.built_map, sd.did_update) != (so.had_points, so.built_map, so.did_update):
            rec["n_iters_equal"] = False
        # ---- the pipeline from raw inputs
        st_d, st_o = lm.get_state(), lio.get_state()
        pe = pose_rel_err(st_d, st_o)
        row["pose_rel_err"] = pe
        rec["max_pose_rel_err"] = max(rec["max_pose_rel_err"], pe)
        if so.had_points:
            down_d = lm.device.scan_get_down(max(sd.n_down, 1))
            down_o = lio.feats_down()
            kd, ko = voxel_keys(down_d), voxel_keys(down_o[:, 0:3])
            vdiff = len(np.setxor1d(kd, ko))
            row["voxels_differing"] = vdiff
            rec["voxels_differing"] += vdiff
            rec["voxels_total"] += so.n_down
        if so.did_update:
            if sd.n_iters != so.n_iters:
                rec["n_iters_equal"] = False
            its_d, its_o = lm.iters(), lio.iters()
            for a, b in zip(its_d, its_o):
                if (a.did_match, a.ekf_stop, a.converged) != (b.did_match, b.ekf_stop, b.converged):
                    rec["n_iters_equal"] = False
                rec["pipeline_effct_max_abs_diff"] = max(rec["pipeline_effct_max_abs_diff"], abs(a.effct_feat_num - b.effct_feat_num))
                ipe = pose_rel_err(np.array(a.state_out), np.array(b.state_out))
                rec["max_iter_pose_rel_err"] = max(rec["max_iter_pose_rel_err"], ipe)
            if its_d and its_o:
                row["effct_dev"], row["effct_orc"] = its_d[-1].effct_feat_num, its_o[-1].effct_feat_num
            rec["pipeline_added_max_abs_diff"] = max(rec["pipeline_added_max_abs_diff"], abs(sd.added - so.added))
            row["added_dev"], row["added_orc"] = sd.added, so.added
        # ---- identical inputs: the oracle's feats_down, poses and map-before through dlt_measure
        if dm is not None and so.did_update:
            down = lio.feats_down()
            dm.scan_set_down(np.column_stack([down[:, 0:3], down[:, 8]]).astype(np.float32))
            for it in lio.iters():
                m = dm.measure(np.array(it.pose_in), bool(it.did_match))
                if m.effct_feat_num != it.effct_feat_num or m.n_down != so.n_down:
                    rec["effct_equal"] = False
                if it.effct_feat_num > 0:
                    rec["max_HtH_rel_err"] = max(rec["max_HtH_rel_err"], rel_err(m.HtH, np.array(it.HtH).reshape(12, 12)))
                    rec["max_Htr_rel_err"] = max(rec["max_Htr_rel_err"], rel_err(m.Htr, np.array(it.Htr)))
            near_o, d2_o, cnt_o, sel_o = lio.nearest()
            nbr_d, cnt_d, sel_d = dm.get_nearest(so.n_down)
            same = np.array_equal(cnt_d, cnt_o) and np.array_equal(nbr_d[:, :, :3].view(np.uint32), near_o[:, :, :3].view(np.uint32)) and \
                np.array_equal(nbr_d[:, :, 3].view(np.uint32), d2_o.view(np.uint32))
            rec["knn_queries"] += int(so.n_down)
            if not same:
                rec["knn_sets_equal"] = False
                row["knn_rows_differing"] = int((np.abs(nbr_d[:, :, :3] - near_o[:, :, :3]).reshape(len(cnt_o), -1).max(1) > 0).sum())
            if not np.array_equal(sel_d, sel_o):
                rec["selected_equal"] = False
            if not so.ekf_stop:
                n_ds, n_raw = dm.map_incremental(st_o[:24], True)
                if (n_ds, n_raw) != (so.n_added_ds, so.n_added_raw):
                    rec["add_lists_equal"] = False
        # ---- map contents
        if self.check_maps and so.had_points:
            flat = lio.map().flatten()
            if dm is not None:
                d = map_set_diff(dm.map_export(), flat)
                if d:
                    rec["map_contents_equal"] = False
                    row["identical_input_map_diff"] = d
            exp = lm.device.map_export()
            d = map_set_diff(exp, flat)
            row["pipeline_map_points_differing_bitwise"] = d  # (inserted centroids differ in the last bit: fixed-point vs float sums)
            dq = map_quantised_diff(exp, flat) if d else 0
            row["pipeline_map_points_differing_1mm"] = dq
            rec["pipeline_map_points_differing_max"] = max(rec["pipeline_map_points_differing_max"], dq)
            if d and not self.chained:  # keep "identical inputs" true for the next scan
                lm.device.map_build(flat)
                rec["resyncs"] += 1
        self.series.append(row)
        return so, sd

    def summary(self) -> dict:
        return dict(self.rec)
