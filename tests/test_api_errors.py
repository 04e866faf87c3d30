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
    # a bounding box whose voxel index space exceeds the occupancy bitmap
    small = ScanToMap(lib, max_scan_points=256, max_map_points=4096, voxel_bitmap_bits=1 << 12)
    far = np.zeros((8, 12), np.float32)
    far[:, 0] = np.linspace(-400, 400, 8)
    far[:, 1] = np.linspace(-400, 400, 8)
    small.scan_deskew(far)
    with pytest.raises(DltError) as e:
        small.scan_downsample()
    assert "capacity" in code(e)
    small.close()
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
