"""Error behaviour of the C ABI: bad arguments, call-sequence errors and exceeded capacities come back as status codes
with a message (nothing throws across the boundary, nothing falls back to a CPU path), on both backends."""
import ctypes as C

import numpy as np
import pytest

from daliti_b200.binding import DltError, ScanToMap


def code(exc):
    return str(exc.value).split(":")[0]


def test_call_sequence_errors(dev):
    lib, _ = dev
    dm = ScanToMap(lib, max_scan_points=1024, max_map_points=4096)
    with pytest.raises(DltError) as e:
        dm.scan_downsample()                      # no scan yet
    assert "call-sequence" in code(e)
    with pytest.raises(DltError) as e:
        dm.measure(np.eye(3).ravel().tolist() + [0] * 15, True)  # no downsampled scan
    assert "call-sequence" in code(e)
    dm.scan_set_down(np.zeros((4, 4), np.float32))
    with pytest.raises(DltError) as e:
        dm.measure(np.eye(3).ravel().tolist() + [0] * 15, False)  # do_match = 0 before any match pass
    assert "call-sequence" in code(e)
    # dlt_iekf_update: iteration count outside [1, DLT_IEKF_MAX_ITER]
    blk = (C.c_double * 8192)()                    # larger than dlt_iekf_block; all zero -> max_iteration = 0
    rc = lib.dlt_iekf_update(dm.h, blk, None, None, None)
    assert rc != 0 and b"max_iteration" in lib.dlt_last_error(dm.h)
    dm.close()


def test_capacity_errors(dev):
    lib, _ = dev
    dm = ScanToMap(lib, max_scan_points=256, max_map_points=4096)
    pts = np.zeros((300, 12), np.float32)
    with pytest.raises(DltError) as e:
        dm.scan_deskew(pts)                        # scan larger than max_scan_points
    assert "capacity" in code(e)
    with pytest.raises(DltError) as e:
        dm.scan_set_down(np.zeros((300, 4), np.float32))
    assert "capacity" in code(e)
    # front end: bad layouts
    cloud = np.zeros(64, np.uint8)
    with pytest.raises(DltError):
        dm.frontend_sample(cloud, (30, 0, 4, 8, 16, 20, 24), "velodyne")        # point_step not a multiple of 4
    with pytest.raises(DltError):
        dm.frontend_sample(cloud, (32, 0, 4, 8, 16, 20, 30), "velodyne")        # time field runs past the record
    with pytest.raises(DltError):
        dm.frontend_sample(np.zeros(48 * 400, np.uint8), (48, 0, 4, 8, 16, 26, 20), "ouster", point_filter_num=1)  # 400 > max_scan_points
    dm.close()


def test_empty_inputs(dev):
    lib, _ = dev
    dm = ScanToMap(lib, max_scan_points=1024, max_map_points=4096)
    dm.scan_deskew(np.zeros((0, 12), np.float32))
    assert dm.scan_downsample() == 0
    m = dm.measure(np.eye(3).ravel().tolist() + [0] * 3 + np.eye(3).ravel().tolist() + [0] * 3, True)  # empty scan, empty map
    assert m.effct_feat_num == 0 and m.n_down == 0 and not m.HtH.any()
    assert dm.map_valid_count() == 0
    assert dm.map_delete_boxes(np.array([[-1, -1, -1, 1, 1, 1]], np.float32)) == 0
    assert len(dm.map_export()) == 0
    ptr, n, ts, span, sh = dm.frontend_sample(np.zeros(0, np.uint8), (32, 0, 4, 8, 16, 20, 24), "velodyne")
    assert n == 0
    dm.close()


def test_voxelgrid_bitmap_grows_instead_of_failing(dev):
    """A bounding box whose voxel-index space exceeds the occupancy bitmap (one far outlier, a small leaf at long range):
    pcl::VoxelGrid handles anything up to INT_MAX cells, so the bitmap grows to what the scan needs and the result equals
    that of a handle whose bitmap was large enough from the start -- on the synchronous call and on the first evaluation
    behind dlt_scan_downsample_async."""
    lib, _ = dev
    rng = np.random.default_rng(3)
    pts = np.zeros((200, 12), np.float32)
    pts[:, 0:3] = rng.uniform(-20, 20, (200, 3))
    pts[0, 0:3] = [-400, -390, 5]   # outliers: 1600 x 1600 x ~80 cells at the 0.5 m leaf
    pts[1, 0:3] = [400, 395, -3]
    pts[:, 8] = rng.uniform(0, 100, 200)
    big = ScanToMap(lib, max_scan_points=256, max_map_points=4096)
    big.scan_deskew(pts)
    n_big = big.scan_downsample()
    ref = big.scan_get_down(256)
    small = ScanToMap(lib, max_scan_points=256, max_map_points=4096, voxel_bitmap_bits=1 << 12)
    small.scan_deskew(pts)
    assert small.scan_downsample() == n_big
    assert np.array_equal(small.scan_get_down(256), ref)
    small.close()
    small = ScanToMap(lib, max_scan_points=256, max_map_points=4096, voxel_bitmap_bits=1 << 12)
    small.map_build(np.column_stack([rng.uniform(-20, 20, (500, 3)), np.zeros(500)]).astype(np.float32))
    small.scan_deskew(pts)
    assert lib.dlt_scan_downsample_async(small.h) == 0
    pose = np.eye(3).ravel().tolist() + [0] * 3 + np.eye(3).ravel().tolist() + [0] * 3
    m = small.measure(pose, True)   # finds the overflow in the result block, grows, runs the VoxelGrid and itself again
    assert m.n_down == n_big
    assert np.array_equal(small.scan_get_down(256), ref)
    big.scan_set_down(ref)
    big.map_build(small.map_export())
    m2 = big.measure(pose, True)
    assert m.effct_feat_num == m2.effct_feat_num and np.array_equal(m.HtH, m2.HtH)
    small.close()
    big.close()


def test_downsample_twice_is_idempotent(dev):
    """k_vox_final resets the bounding box for the next scan; a second dlt_scan_downsample of the SAME scan (async followed by
    sync, or a retry) has to recompute it instead of decoding the reset pattern."""
    lib, _ = dev
    rng = np.random.default_rng(4)
    pts = np.zeros((1000, 12), np.float32)
    pts[:, 0:3] = rng.uniform(-15, 15, (1000, 3))
    dm = ScanToMap(lib, max_scan_points=1024, max_map_points=4096)
    dm.scan_deskew(pts)
    n1 = dm.scan_downsample()
    d1 = dm.scan_get_down(1024)
    v1 = dm.scan_get_voxel_of_point(1000)
    n2 = dm.scan_downsample()
    assert n2 == n1 and np.array_equal(dm.scan_get_down(1024), d1) and np.array_equal(dm.scan_get_voxel_of_point(1000), v1)
    assert lib.dlt_scan_downsample_async(dm.h) == 0
    assert dm.scan_downsample() == n1 and np.array_equal(dm.scan_get_down(1024), d1)
    dm.close()


def test_map_overflow_leaves_a_safe_handle(dev):
    """After the bucket pool is exhausted the device counter runs past the pool: host-side loops over the buckets must clamp
    it, mutation and updates are refused until dlt_map_build starts over, reads still work."""
    lib, _ = dev
    dm = ScanToMap(lib, max_scan_points=8192, max_map_points=4096)   # 4096 buckets (the minimum pool)
    rng = np.random.default_rng(5)
    spread = np.column_stack([rng.uniform(-400, 400, (8000, 3)), np.zeros(8000)]).astype(np.float32)  # ~8000 distinct cells
    with pytest.raises(DltError) as e:
        dm.map_build(spread)
    assert "capacity" in code(e)
    exported = dm.map_export()            # clamped: no out-of-bounds read
    assert len(exported) <= 4096 * 7
    with pytest.raises(DltError) as e:
        dm.map_delete_boxes(np.array([[-500, -500, -500, 500, 500, 500]], np.float32))
    assert "capacity" in code(e) and b"overflowed" in lib.dlt_last_error(dm.h)
    with pytest.raises(DltError):
        dm.map_add(spread[:10], False)
    dm.scan_set_down(spread[:100])
    with pytest.raises(DltError):
        dm.measure(np.eye(3).ravel().tolist() + [0] * 3 + np.eye(3).ravel().tolist() + [0] * 3, True)
    dm.map_build(spread[:1000])           # a fresh map revives the handle
    assert dm.map_valid_count() == 1000
    assert dm.map_delete_boxes(np.array([[-500, -500, -500, 500, 500, 500]], np.float32)) == 1000
    dm.close()


def test_shard_tile_shift_below_the_halo_is_rejected(dev):
    """shard_keeps_cell tests the corners of the +-4-cell halo cube, which is only sufficient for tiles of >= 8 cells."""
    lib, _ = dev
    for shift, ok in ((1, False), (2, False), (3, True), (0, True)):
        try:
            dm = ScanToMap(lib, max_scan_points=256, max_map_points=4096, shard_rank=0, shard_count=2, shard_tile_shift=shift)
            dm.close()
            assert ok, shift
        except DltError:
            assert not ok, shift
