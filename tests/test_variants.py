"""Opt-in kernel variants (daliti_b200/csrc/Makefile `variants`) must be result-identical to the default build: the same
kernel sources compiled with the variant's macro under the kernel-logic emulator, compared bit for bit."""
import os
import subprocess

import numpy as np

import helpers
from daliti_b200 import synth
from daliti_b200.binding import ScanToMap, load_library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_emu_variant(tmp_path, *defines):
    emu = os.path.join(ROOT, "tests", "emu")
    csrc = os.path.join(ROOT, "daliti_b200", "csrc")
    out = str(tmp_path / "libdaliti_emu_variant.so")
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-DDLT_EMU", *["-D" + d for d in defines], "-I" + emu, "-I" + csrc, "-w", "-shared",
           "-x", "c++", os.path.join(csrc, "dlt_api.cu"), os.path.join(csrc, "host", "eskf_lio_host.cpp"), "-x", "c++", os.path.join(emu, "cuda_emu.cpp"), "-o", out]
    subprocess.run(cmd, check=True)
    return load_library(out)


def _match_pass(lib, map_pts, down, pose):
    dm = ScanToMap(lib, max_scan_points=8192, max_map_points=1 << 17)
    dm.map_build(map_pts)
    dm.scan_set_down(down)
    m = dm.measure(pose, True)
    first = (m.effct_feat_num, m.n_unresolved, m.HtH.copy(), m.Htr.copy())
    nbr, cnt, sel = dm.get_nearest(len(down))
    n_ds, n_raw = dm.map_incremental(pose, True)
    out = (first, nbr.copy(), cnt.copy(), sel.copy(), (n_ds, n_raw), dm.map_export())
    dm.close()
    return out


def test_knn8_prune_variant_is_result_identical(emu_lib, tmp_path):
    """DLT_KNN8_PRUNE: candidates / cells beyond the radius that can finish a query in the first pass are skipped"""
    var = _build_emu_variant(tmp_path, "DLT_KNN8_PRUNE=1")
    for seed, sparse in ((31, False), (32, True)):
        seq = helpers.small_sequence(seed=seed, half=30.0, beams=16, azimuths=240, n_boxes=8)
        map_pts = synth.sample_map(seq.scene, seed=seed)
        if sparse:  # a map that covers part of the scene: many unresolved / far queries
            map_pts = map_pts[map_pts[:, 0] < 0.0]
        pts, t_beg, imu = seq.scan(0)
        dm = ScanToMap(emu_lib, max_scan_points=8192, max_map_points=4096)
        dm.scan_deskew(pts)
        down = dm.scan_get_down(dm.scan_downsample())
        dm.close()
        pose = seq.traj.pose24(0.1)
        a = _match_pass(emu_lib, map_pts, down, pose)
        b = _match_pass(var, map_pts, down, pose)
        assert a[0][0] == b[0][0] and a[0][1] == b[0][1]
        np.testing.assert_array_equal(a[0][2], b[0][2])  # H^T H: same points, same order of summation
        np.testing.assert_array_equal(a[0][3], b[0][3])
        for x, y in zip(a[1:4], b[1:4]):
            np.testing.assert_array_equal(x, y)  # neighbours (coordinates, d2), counts, point_selected_surf
        assert a[4] == b[4]
        assert set(map(tuple, a[5].tolist())) == set(map(tuple, b[5].tolist()))


def _replay(lib, n_scans=3):
    from daliti_b200.lio import LaserMapping

    seq = helpers.small_sequence(seed=41, half=30.0, beams=16, azimuths=240, n_boxes=8)
    map_pts = synth.sample_map(seq.scene, seed=41)
    lm = LaserMapping(lib, dev=dict(max_scan_points=8192, max_map_points=1 << 17), featptsThreshold=5, device_loop=0)
    lm.force_imu_ready([0.0, 0.0, synth.G], np.concatenate([[seq.t_start - 0.005], [0, 0, synth.G], [0, 0, seq.traj.yaw_rate]]))
    lm.set_state(helpers.state612(seq.traj, seq.t_start))
    lm.device.map_build(map_pts)
    out = []
    for k in range(n_scans):
        pts, t_beg, imu = seq.scan(k)
        lm.on_lidar_msg()
        o = lm.process_scan(pts, t_beg, imu)
        out.append((lm.get_state().copy(), o.n_iters, o.added, [np.array(it.HtH) for it in lm.iters()]))
    lm.close()
    return out


def test_zerocopy_result_path_is_result_identical(emu_lib, monkeypatch):
    """DLT_ZEROCOPY=1: k_residual's last block stores the result block into pinned host memory and raises a flag the host
    spins on, instead of a device->host copy + stream synchronisation per iteration (host loop).  Same numbers."""
    a = _replay(emu_lib)
    monkeypatch.setenv("DLT_ZEROCOPY", "1")
    b = _replay(emu_lib)
    for (sa, ia, aa, ha), (sb, ib, ab, hb) in zip(a, b):
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
