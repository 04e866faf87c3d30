"""LiDAR front end with feature_enabled = 0: FeatureExtract::cachePointCloud + samplePointCloud
(eskf_lio/src/feature_extract.cpp:264-450) on the device vs the oracle's restatement, for the four sensor record layouts
of eskf_lio/include/my_utility.h:19-54.  Byte-exact: the kept records, their order and every field."""
import numpy as np
import pytest

SENSORS = {"velodyne": 0, "livox": 1, "ouster": 2, "robosense": 3}


def make_cloud(sensor, n, seed):
    """a synthetic PointCloud2 payload in the sensor's own record layout -> (uint8 data, layout)"""
    rng = np.random.default_rng(seed)
    xyz = rng.normal(0, 12.0, (n, 3)).astype(np.float32)
    xyz[rng.random(n) < 0.03] *= 0.01          # inside lidarMinRange
    xyz[rng.random(n) < 0.02] *= 40.0          # beyond lidarMaxRange
    ring = rng.integers(0, 64, n).astype(np.uint16)
    frac = np.sort(rng.random(n))
    if sensor == "velodyne":    # VelodynePointXYZIRT: x y z pad | intensity ring(u16) pad time(f32) pad
        dt = np.dtype({"names": ["x", "y", "z", "intensity", "ring", "time"], "formats": ["<f4", "<f4", "<f4", "<f4", "<u2", "<f4"],
                       "offsets": [0, 4, 8, 16, 20, 24], "itemsize": 32})
        a = np.zeros(n, dt)
        a["time"] = (frac * 0.1 - 0.1).astype(np.float32)   # LIO-SAM 6-axis data: times relative to the sweep end
        a["intensity"] = rng.uniform(0, 255, n).astype(np.float32)
    elif sensor == "livox":
        dt = np.dtype({"names": ["x", "y", "z", "intensity", "ring", "time"], "formats": ["<f4", "<f4", "<f4", "<f4", "<u2", "<f4"],
                       "offsets": [0, 4, 8, 16, 20, 24], "itemsize": 32})
        a = np.zeros(n, dt)
        a["time"] = (frac * 0.1 + 1e-4).astype(np.float32)
        a["intensity"] = rng.uniform(0, 255, n).astype(np.float32)
    elif sensor == "ouster":    # OusterPointXYZIRT: x y z pad | intensity t(u32) reflectivity ring ambient range
        dt = np.dtype({"names": ["x", "y", "z", "intensity", "time", "reflectivity", "ring", "ambient", "range"],
                       "formats": ["<f4", "<f4", "<f4", "<f4", "<u4", "<u2", "<u2", "<u2", "<u4"], "offsets": [0, 4, 8, 16, 20, 24, 26, 28, 32], "itemsize": 48})
        a = np.zeros(n, dt)
        a["time"] = (frac * 1e8).astype(np.uint32)
        a["intensity"] = rng.uniform(0, 2000, n).astype(np.float32)
    else:                       # rsPointXYZIRT: x y z pad | intensity(u8) ring(u16) timestamp(f64)
        dt = np.dtype({"names": ["x", "y", "z", "intensity", "ring", "time"], "formats": ["<f4", "<f4", "<f4", "u1", "<u2", "<f8"],
                       "offsets": [0, 4, 8, 16, 18, 24], "itemsize": 32})
        a = np.zeros(n, dt)
        a["time"] = 1.7e9 + frac * 0.1
        a["intensity"] = rng.integers(0, 255, n).astype(np.uint8)
        bad = rng.random(n) < 0.05                 # non-dense cloud: NaN returns are dropped before the 1-in-N sampling
        bad[0] = bad[-1] = False
        xyz[bad, rng.integers(0, 3, int(bad.sum()))] = np.nan
    a["x"], a["y"], a["z"], a["ring"] = xyz[:, 0], xyz[:, 1], xyz[:, 2], ring
    lay = (dt.itemsize, dt.fields["x"][1], dt.fields["y"][1], dt.fields["z"][1], dt.fields["intensity"][1], dt.fields["ring"][1], dt.fields["time"][1])
    return a.view(np.uint8).reshape(-1).copy(), lay


@pytest.mark.parametrize("sensor", ["velodyne", "livox", "ouster", "robosense"])
@pytest.mark.parametrize("filter_num", [1, 5])
def test_frontend_sample_byte_exact(dev, oracle, sensor, filter_num):
    from daliti_b200.binding import ScanToMap

    lib, is_gpu = dev
    n = 65536 if is_gpu else 6000
    dm = ScanToMap(lib, max_scan_points=n, max_map_points=4096)
    for seed, count in ((1, n), (2, n - 3), (3, 7)):
        data, lay = make_cloud(sensor, count, seed)
        want, ts_o, sh_o = oracle.frontend_sample(data, lay, SENSORS[sensor], filter_num, 0.5, 200.0)
        ptr, m, ts, span, sh = dm.frontend_sample(data, lay, sensor, filter_num, 0.5, 200.0)
        assert m == len(want), (sensor, filter_num, seed)
        assert ts == ts_o and sh == sh_o
        got = dm.frontend_read(m)
        assert got.tobytes() == want.tobytes()
        if m:
            assert span == float(want[-1, 6])  # observation_end_time = lidar_beg_time + points.back().normal_z (laserMapping.cpp:546)
            assert 0 < m <= (count + filter_num - 1) // filter_num
    dm.close()


def test_frontend_feeds_the_update(dev, oracle):
    """the sampled records stay on the device and go straight into deskew -> VoxelGrid: same downsampled cloud as handing the
    oracle's /laser_cloud_surf records over from the host"""
    from daliti_b200.binding import ScanToMap

    lib, is_gpu = dev
    data, lay = make_cloud("ouster", 20000 if is_gpu else 4000, 9)
    want, _, _ = oracle.frontend_sample(data, lay, SENSORS["ouster"], 5, 0.5, 200.0)
    a = ScanToMap(lib, max_scan_points=32768, max_map_points=4096)
    b = ScanToMap(lib, max_scan_points=32768, max_map_points=4096)
    ptr, m, _, _, _ = a.frontend_sample(data, lay, "ouster", 5, 0.5, 200.0)
    a._ck(a.lib.dlt_scan_deskew_dev(a.h, __import__("ctypes").c_void_p(ptr), m, None, 0, None))
    b.scan_deskew(want)
    na, nb = a.scan_downsample(), b.scan_downsample()
    assert na == nb and na > 0
    assert a.scan_get_down(na).tobytes() == b.scan_get_down(nb).tobytes()
    a.close()
    b.close()


def test_process_cloud_equals_process_scan(dev, oracle):
    """dlt_lio_process_cloud (sensor records -> front end on the device -> update) == dlt_lio_process_scan on the records the
    reference's front end would have published, over a short replay with inserts"""
    import helpers
    from daliti_b200 import synth
    from daliti_b200.lio import LaserMapping

    lib, is_gpu = dev
    seq = helpers.small_sequence(seed=31, half=25.0, beams=16, azimuths=900 if is_gpu else 240, n_boxes=8, speed=2.0, yaw_rate=0.2)
    map_pts = synth.sample_map(seq.scene, seed=31)
    lms = []
    for _ in range(2):
        lm = LaserMapping(lib, dev=dict(max_scan_points=32768, max_map_points=4 * len(map_pts)), featptsThreshold=5)
        lm.force_imu_ready([0.0, 0.0, synth.G], np.concatenate([[seq.t_start - 0.005], [0, 0, synth.G], [0, 0, seq.traj.yaw_rate]]))
        lm.set_state(helpers.state612(seq.traj, seq.t_start))
        lm.device.map_build(map_pts)
        lms.append(lm)
    dt = np.dtype({"names": ["x", "y", "z", "intensity", "time", "reflectivity", "ring", "ambient", "range"],
                   "formats": ["<f4", "<f4", "<f4", "<f4", "<u4", "<u2", "<u2", "<u2", "<u4"], "offsets": [0, 4, 8, 16, 20, 24, 26, 28, 32], "itemsize": 48})
    lay = (48, 0, 4, 8, 16, 26, 20)
    for k in range(3):
        pts, t_beg, imu = seq.scan(k)
        # an Ouster message carrying this sweep: t in ns from the sweep start, ring, intensity
        a = np.zeros(len(pts), dt)
        a["x"], a["y"], a["z"], a["intensity"] = pts[:, 0], pts[:, 1], pts[:, 2], pts[:, 8]
        a["ring"] = pts[:, 5].astype(np.uint16)
        a["time"] = np.round(pts[:, 4].astype(np.float64) * 0.1 * 1e9).astype(np.uint32)
        data = a.view(np.uint8).reshape(-1)
        want, _, _ = oracle.frontend_sample(data, lay, SENSORS["ouster"], 3, 0.5, 200.0)
        for lm in lms:
            lm.on_lidar_msg()
        oa, ns = lms[0].process_cloud(data, lay, "ouster", t_beg, imu, point_filter_num=3, min_range=0.5, max_range=200.0)
        ra = (oa.n_raw, oa.n_down, oa.n_iters, oa.added, oa.ekf_stop)
        ob = lms[1].process_scan(want, t_beg, imu)
        rb = (ob.n_raw, ob.n_down, ob.n_iters, ob.added, ob.ekf_stop)
        assert ns == len(want) and ra == rb, (k, ns, len(want), ra, rb)
        np.testing.assert_array_equal(lms[0].get_state(), lms[1].get_state())
    for lm in lms:
        lm.close()
