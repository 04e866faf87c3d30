"""The measurement model (laserMapping.cpp:829-979) and map_incremental (:582-630) on the
device vs the restated reference loop, iteration by iteration, on the oracle's own inputs:
same downsampled scan, same pose, same map contents.
  * neighbour sets and point_selected_surf: bit-exact
  * effct_feat_num: exact;  H^T H, H^T r, residual sum: rel 1e-10 (fp64, summation order differs)
  * map contents after map_incremental: equal as sets."""
import numpy as np
import pytest

import helpers
from daliti_b200 import synth
from daliti_b200.binding import ScanToMap
from oracle_binding import MAP_PORT, MAP_REF


def as_set(a):
    return set(map(tuple, np.asarray(a)[:, :3].tolist()))


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def run_sequence(lib, oracle, seq, map_pts, n_scans, ext=False, max_pts=4096, check_map=True, read_nearest=True, pre_delete=None, **cfg_kw):
    kind = MAP_REF if oracle.ref_ok else MAP_PORT
    lio = helpers.start_oracle_lio(oracle, seq, map_pts, kind, extrinsic_est_en=1 if ext else 0, **cfg_kw)
    dm = ScanToMap(lib, max_scan_points=max_pts, max_map_points=max(4 * len(map_pts), 16384), extrinsic_est_en=1 if ext else 0)
    dm.map_build(map_pts)
    if pre_delete is not None:  # boxes removed from both maps before the first scan (Delete_Point_Boxes, ikd_Tree.cpp:631-658)
        boxes = np.asarray(pre_delete, np.float32).reshape(-1, 6)
        assert dm.map_delete_boxes(boxes) == lio.map().delete_boxes(boxes) > 0
    total_eff = 0
    for k in range(n_scans):
        pts, t_beg, imu = seq.scan(k)
        lio.on_lidar_msg()
        # the device map must hold what the oracle's map holds BEFORE this scan (delete boxes included)
        s = lio.process_scan(pts, t_beg, imu)
        assert s.had_points and s.did_update
        down = lio.feats_down()
        xyzi = np.column_stack([down[:, 0:3], down[:, 8]]).astype(np.float32)
        dm.scan_set_down(xyzi)
        its = lio.iters()
        assert len(its) >= 1
        for it in its:
            m = dm.measure(np.array(it.pose_in), bool(it.did_match))
            assert m.n_down == s.n_down
            assert m.effct_feat_num == it.effct_feat_num, (k, it.iter)
            total_eff += it.effct_feat_num
            if it.effct_feat_num > 0:
                assert rel_err(m.HtH, np.array(it.HtH).reshape(12, 12)) < 1e-10
                assert rel_err(m.Htr, np.array(it.Htr)) < 1e-9
                assert abs(m.total_residual - it.total_residual) <= 1e-10 * max(1.0, it.total_residual)
            # degradation output: device eigen-decomposition of the 6x6 pose block (property-checked: the
            # reference has no counterpart) -- eigenvalues vs LAPACK, orthonormal vectors, reconstruction
            evals, evecs = dm.degeneracy()
            A6 = m.HtH[:6, :6]
            ev = np.linalg.eigvalsh(A6)
            scale = max(1.0, abs(ev).max())
            np.testing.assert_allclose(evals, ev, rtol=1e-9, atol=1e-9 * scale)
            np.testing.assert_allclose(evecs.T @ evecs, np.eye(6), atol=1e-10)
            np.testing.assert_allclose(evecs @ np.diag(evals) @ evecs.T, A6, atol=1e-9 * scale)
        # neighbours of the last match pass + point_selected_surf after the last iteration
        if read_nearest:  # (reading them makes every neighbour set exact first: the full fallback)
            near_o, d2_o, cnt_o, sel_o = lio.nearest()
            nbr_d, cnt_d, sel_d = dm.get_nearest(s.n_down)
            np.testing.assert_array_equal(cnt_d, cnt_o)
            np.testing.assert_array_equal(nbr_d[:, :, :3], near_o[:, :, :3])
            np.testing.assert_array_equal(nbr_d[:, :, 3], d2_o)
            np.testing.assert_array_equal(sel_d, sel_o)
        # map_incremental with the oracle's post-update state
        st = lio.get_state()
        if not s.ekf_stop:
            n_ds, n_raw = dm.map_incremental(st[:24], True)
            assert (n_ds, n_raw) == (s.n_added_ds, s.n_added_raw)
        if check_map:
            assert dm.map_valid_count() == lio.map().validnum()
            assert as_set(dm.map_export()) == as_set(lio.map().flatten())
    dm.close()
    return total_eff


def test_measure_box_world(dev, oracle):
    lib, is_gpu = dev
    if is_gpu:
        seq = helpers.small_sequence(seed=1, half=60.0, beams=32, azimuths=1024, n_boxes=24)
        n_scans = 6
    else:
        seq = helpers.small_sequence(seed=1, half=25.0, beams=16, azimuths=240, n_boxes=8)
        n_scans = 3
    map_pts = synth.sample_map(seq.scene, seed=1)
    eff = run_sequence(lib, oracle, seq, map_pts, n_scans, max_pts=65536 if is_gpu else 8192, featptsThreshold=5)
    assert eff > 200


def test_measure_extrinsic_enabled(dev, oracle):
    """extrinsic_est_en = true: the full 12-column Jacobian (laserMapping.cpp:968-972)"""
    lib, is_gpu = dev
    seq = helpers.small_sequence(seed=2, half=25.0, beams=16, azimuths=240 if not is_gpu else 900, n_boxes=8)
    map_pts = synth.sample_map(seq.scene, seed=2)
    eff = run_sequence(lib, oracle, seq, map_pts, 2, ext=True, max_pts=32768 if is_gpu else 8192, featptsThreshold=5)
    assert eff > 100


def test_measure_sparse_map_far_points(dev, oracle):
    """a map that covers only part of the scene: many queries have no neighbour within the search
    rings, exercising the exact fallback that map_incremental depends on (laserMapping.cpp:593-617)"""
    lib, is_gpu = dev
    seq = helpers.small_sequence(seed=3, half=25.0, beams=16, azimuths=240 if not is_gpu else 900, n_boxes=6)
    map_pts = synth.sample_map(seq.scene, seed=3, region=(-8.0, 8.0, -8.0, 8.0))
    run_sequence(lib, oracle, seq, map_pts, 3, max_pts=32768 if is_gpu else 8192, featptsThreshold=5)


def test_map_incremental_far_points_nearest_only(dev, oracle):
    """same sparse map, but map_incremental runs straight after the iterations: unresolved queries are classified from
    what the rings saw plus, for those that saw nothing, the single nearest map point (k_nn1) -- the add lists and
    the map contents must still equal the reference's, which searched its tree exactly"""
    lib, is_gpu = dev
    seq = helpers.small_sequence(seed=3, half=25.0, beams=16, azimuths=240 if not is_gpu else 900, n_boxes=6)
    map_pts = synth.sample_map(seq.scene, seed=3, region=(-8.0, 8.0, -8.0, 8.0))
    run_sequence(lib, oracle, seq, map_pts, 3, max_pts=32768 if is_gpu else 8192, read_nearest=False, featptsThreshold=5)




def test_map_incremental_far_points_after_box_delete(dev, oracle):
    """the nearest-point pass for far queries (k_nn1_seed / k_nn1) works from one LIVE sample per pool run and skips runs whose
    cell box is too far: with part of the map box-deleted the runs there hold dead points only (and boxes that are still
    set), so a bound taken from a dead point, or a run skipped although it holds the nearest live one, would change the add
    lists or the map contents against the reference tree"""
    lib, is_gpu = dev
    seq = helpers.small_sequence(seed=3, half=25.0, beams=16, azimuths=240 if not is_gpu else 900, n_boxes=6)
    map_pts = synth.sample_map(seq.scene, seed=3, region=(-8.0, 8.0, -8.0, 8.0))
    boxes = [[-8.5, -8.5, -5.0, -2.0, 8.5, 50.0], [3.0, 2.0, -5.0, 8.5, 8.5, 50.0]]  # two sides of the mapped patch
    run_sequence(lib, oracle, seq, map_pts, 3, max_pts=32768 if is_gpu else 8192, read_nearest=False, pre_delete=boxes, featptsThreshold=5)
