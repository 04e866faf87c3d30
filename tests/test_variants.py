"""Switchable paths of the library must be result-identical to each other: the zero-copy result block against the copy +
synchronise path, map_incremental off the critical path (async_insert) against the synchronous call."""
import os

import numpy as np
import pytest

import helpers
from daliti_b200 import synth
from daliti_b200.binding import ScanToMap, load_library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _replay(lib, n_scans=3, collect=True, **cfg):
    from daliti_b200.lio import LaserMapping

    seq = helpers.small_sequence(seed=41, half=30.0, beams=16, azimuths=240, n_boxes=8)
    map_pts = synth.sample_map(seq.scene, seed=41)
    lm = LaserMapping(lib, dev=dict(max_scan_points=8192, max_map_points=1 << 17), featptsThreshold=5, device_loop=0, **cfg)
    lm.collect_after_scan = collect
    lm.force_imu_ready([0.0, 0.0, synth.G], np.concatenate([[seq.t_start - 0.005], [0, 0, synth.G], [0, 0, seq.traj.yaw_rate]]))
    lm.set_state(helpers.state612(seq.traj, seq.t_start))
    lm.device.map_build(map_pts)
    out = []
    for k in range(n_scans):
        pts, t_beg, imu = seq.scan(k)
        lm.on_lidar_msg()
        o = lm.process_scan(pts, t_beg, imu)
        out.append((lm.get_state().copy(), o.n_iters, o.added, [np.array(it.HtH) for it in lm.iters()]))
    if not collect:
        out.append(lm.collect_insert())
    out.append(set(map(tuple, lm.device.map_export().tolist())))
    lm.close()
    return out


def test_zerocopy_result_path_is_result_identical(emu_lib, monkeypatch):
    """DLT_ZEROCOPY=1: k_residual's last block stores the result block into pinned host memory and raises a flag the host
    spins on, instead of a device->host copy + stream synchronisation per iteration (host loop).  Same numbers."""
    a = _replay(emu_lib)
    monkeypatch.setenv("DLT_ZEROCOPY", "1")
    b = _replay(emu_lib)
    assert a[-1] == b[-1]  # map contents
    for (sa, ia, aa, ha), (sb, ib, ab, hb) in zip(a[:-1], b[:-1]):
        assert (ia, aa) == (ib, ab)
        np.testing.assert_array_equal(sa, sb)
        for x, y in zip(ha, hb):
            np.testing.assert_array_equal(x, y)
    # the empty scan takes the same road
    dm = ScanToMap(emu_lib, max_scan_points=1024, max_map_points=4096)
    dm.map_build(np.array([[0, 0, 0, 1], [1, 0, 0, 1], [0, 1, 0, 1], [0, 0, 1, 1], [1, 1, 0, 1], [1, 1, 1, 1]], np.float32))
    dm.scan_deskew(np.zeros((0, 12), np.float32))
    assert dm.scan_downsample() == 0
    pose = np.zeros(24)
    pose[[0, 4, 8, 12, 16, 20]] = 1.0
    m = dm.measure(pose, True)
    assert m.effct_feat_num == 0 and not m.HtH.any()
    dm.close()


def _check_async_insert(lib, n_scans):
    """async_insert = 1 (library default): map_incremental runs on its own stream and process_scan returns with added = -1;
    the next scan's deskew / VoxelGrid overlap it and its first match pass waits for it on the device.  Against
    async_insert = 0: the same states, iterations, normal equations, add counts and map contents, bit for bit."""
    sync = _replay(lib, n_scans, async_insert=0)
    waited = _replay(lib, n_scans, async_insert=1)            # the wrapper collects after every scan
    free = _replay(lib, n_scans, collect=False, async_insert=1)  # nothing waits on the host until the end
    assert sync[-1] == waited[-1] == free[-1]
    for k in range(n_scans):
        (s0, i0, a0, h0), (s1, i1, a1, h1), (s2, i2, a2, h2) = sync[k], waited[k], free[k]
        assert (i0, a0) == (i1, a1) and i0 == i2 and a2 == -1
        np.testing.assert_array_equal(s0, s1)
        np.testing.assert_array_equal(s0, s2)
        for x, y, z in zip(h0, h1, h2):
            np.testing.assert_array_equal(x, y)
            np.testing.assert_array_equal(x, z)
    n_ds, n_raw = free[-2]
    assert n_ds + n_raw == sync[n_scans - 1][2]


def test_async_insert_is_result_identical(emu_lib):
    _check_async_insert(emu_lib, 3)


@pytest.mark.gpu
def test_async_insert_is_result_identical_gpu(gpu_lib):
    _check_async_insert(gpu_lib, 8)


def _check_knn_reuse(lib, monkeypatch, n_scans):
    """Rematch passes prove most neighbour sets unchanged instead of searching again (knn_try_reuse); with DLT_KNN_REUSE=0 every
    match pass searches.  Same states, iterations, normal equations, add counts and map contents, bit for bit."""
    a = _replay(lib, n_scans)
    assert any(len(h) >= 3 for _, _, _, h in a[:n_scans])  # scans with a rematch pass behind the first one
    monkeypatch.setenv("DLT_KNN_REUSE", "0")
    b = _replay(lib, n_scans)
    assert a[-1] == b[-1]
    for (sa, ia, aa, ha), (sb, ib, ab, hb) in zip(a[:n_scans], b[:n_scans]):
        assert (ia, aa) == (ib, ab)
        np.testing.assert_array_equal(sa, sb)
        for x, y in zip(ha, hb):
            np.testing.assert_array_equal(x, y)


def test_knn_reuse_is_result_identical(emu_lib, monkeypatch):
    _check_knn_reuse(emu_lib, monkeypatch, 3)


@pytest.mark.gpu
def test_knn_reuse_is_result_identical_gpu(gpu_lib, monkeypatch):
    _check_knn_reuse(gpu_lib, monkeypatch, 8)


def test_reuse_statistics_and_debug_counters(emu_lib):
    """dlt_debug_counters: the rematch passes of a replay report how many queries they saw and how many of those had to be searched
    again; on a map that covers the scene most neighbour sets are proven unchanged"""
    from daliti_b200.lio import LaserMapping

    seq = helpers.small_sequence(seed=41, half=30.0, beams=16, azimuths=240, n_boxes=8)
    map_pts = synth.sample_map(seq.scene, seed=41)
    lm = LaserMapping(emu_lib, dev=dict(max_scan_points=8192, max_map_points=1 << 17), featptsThreshold=5, device_loop=0)
    lm.force_imu_ready([0.0, 0.0, synth.G], np.concatenate([[seq.t_start - 0.005], [0, 0, synth.G], [0, 0, seq.traj.yaw_rate]]))
    lm.set_state(helpers.state612(seq.traj, seq.t_start))
    lm.device.map_build(map_pts)
    n_rematch = 0
    for k in range(3):
        pts, t_beg, imu = seq.scan(k)
        lm.on_lidar_msg()
        o = lm.process_scan(pts, t_beg, imu)
        n_rematch += sum(1 for it in lm.iters()[1:] if it.did_match) * o.n_down
    c = lm.device.debug_counters()
    assert c[1] == lm.device.map_valid_count() and c[2] == 0
    assert n_rematch > 0 and c[12] == n_rematch  # every query of every rematch pass went through the reuse kernel
    assert 0 <= c[13] < 0.5 * c[12]               # ... and most were not searched again
    lm.close()
