"""Scan preparation vs the oracle: esti_plane inputs aside, this covers
  * per-point deskew  (ImuProcess::UndistortPcl backward pass, IMU_Processing.hpp:332-370)
  * VoxelGrid         (PCL 1.10 semantics, call site laserMapping.cpp:775-776):
    voxel assignment bit-exact, output order = ascending voxel idx, centroids within 1e-6."""
import numpy as np
import pytest

from daliti_b200 import synth
from daliti_b200.binding import ScanToMap


def random_scan(n, seed, span=40.0):
    r = np.random.default_rng(seed)
    p = np.zeros((n, 12), np.float32)
    p[:, 0] = r.uniform(-span, span, n)
    p[:, 1] = r.uniform(-span, span, n)
    p[:, 2] = r.uniform(-2, 6, n)
    p[:, 3] = 1.0
    p[:, 4] = r.uniform(0, 1, n)
    p[:, 5] = r.integers(0, 16, n)
    p[:, 6] = 0.1
    p[:, 8] = r.uniform(1, 100, n)
    return p


def test_voxelgrid_assignment_and_centroids(dev, oracle):
    lib, is_gpu = dev
    n = 131072 if is_gpu else 5000
    pts = random_scan(n, 7, span=60.0 if is_gpu else 12.0)
    pts[: n // 4, :3] *= 0.05  # a dense clump: many points per voxel
    out_o, vop_o, _ = oracle.voxel_grid(pts, 0.5, stable=False)
    dm = ScanToMap(lib, max_scan_points=max(n, 1024), max_map_points=4096)
    dm.scan_deskew(pts)  # no IMU poses: copy-through
    nd = dm.scan_downsample()
    assert nd == len(out_o)
    vop_d = dm.scan_get_voxel_of_point(n)
    np.testing.assert_array_equal(vop_d, vop_o)  # voxel assignment + output order: bit-exact
    down = dm.scan_get_down(nd)
    # centroid = sum / count; PCL's float summation order inside a voxel is unspecified (unstable sort),
    # the device sums in 2^-24 fixed point -> relative tolerance 1e-6 (absolute 1e-5 near zero)
    np.testing.assert_allclose(down[:, :3], out_o[:, :3], rtol=1e-6, atol=2e-5)
    np.testing.assert_allclose(down[:, 3], out_o[:, 8], rtol=1e-5, atol=1e-4)
    # run-to-run reproducibility of the device path (order-independent sums)
    dm.scan_deskew(pts[::-1].copy())
    assert dm.scan_downsample() == nd
    np.testing.assert_array_equal(dm.scan_get_down(nd), down)
    dm.close()


def test_voxelgrid_edge_cases(dev, oracle):
    lib, _ = dev
    dm = ScanToMap(lib, max_scan_points=1024, max_map_points=4096)
    one = random_scan(1, 1)
    dm.scan_deskew(one)
    assert dm.scan_downsample() == 1
    np.testing.assert_allclose(dm.scan_get_down(1)[0, :3], one[0, :3], rtol=1e-6)
    dm.scan_deskew(np.zeros((0, 12), np.float32))
    assert dm.scan_downsample() == 0
    same = np.repeat(random_scan(1, 2), 300, axis=0)
    dm.scan_deskew(same)
    assert dm.scan_downsample() == 1
    np.testing.assert_allclose(dm.scan_get_down(1)[0, :3], same[0, :3], rtol=1e-6)
    neg = random_scan(400, 3)
    neg[:, :3] -= 500.0
    o, vop, _ = oracle.voxel_grid(neg, 0.5)
    dm.scan_deskew(neg)
    assert dm.scan_downsample() == len(o)
    np.testing.assert_array_equal(dm.scan_get_voxel_of_point(400), vop)
    dm.close()


def _poses_and_state(n_pose, seed):
    r = np.random.default_rng(seed)
    traj = synth.Trajectory(speed=2.0, yaw_rate=0.2, z0=1.5)
    ts = np.linspace(0.0, 0.1, n_pose)
    ts[1:] += r.uniform(-0.001, 0.001, n_pose - 1)
    poses = np.zeros((n_pose, 22))
    for i, t in enumerate(ts):
        poses[i, 0] = t
        poses[i, 1:4] = traj.acc(t) + r.normal(0, 0.01, 3)
        poses[i, 4:7] = [0.01, -0.02, traj.yaw_rate]
        poses[i, 7:10] = traj.vel(t)
        poses[i, 10:13] = traj.pos(t)
        poses[i, 13:22] = traj.rot(t).ravel()
    poses[0, 0] = 0.0
    state36 = np.zeros(36)
    state36[0:9] = traj.rot(0.1).ravel()
    state36[9:12] = traj.pos(0.1)
    c, s = np.cos(0.03), np.sin(0.03)
    state36[12:21] = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]]).ravel()
    state36[21:24] = [0.05, -0.02, 0.1]
    return poses, state36


@pytest.mark.parametrize("first_time", [0.0, 0.013])
def test_deskew(dev, oracle, first_time):
    """first_time > first IMU offset exercises the reference's repeated compensation of the begin() point."""
    lib, is_gpu = dev
    n = 131072 if is_gpu else 4000
    pts = random_scan(n, 11)
    pts[:, 4] = np.maximum(pts[:, 4], np.float32(first_time / 0.1 + 1e-4)) if first_time > 0 else pts[:, 4]
    if first_time == 0.0:
        pts[:5, 4] = 0.0  # t == 0 points stay uncompensated (IMU_Processing.hpp:345)
    poses, state36 = _poses_and_state(22, 5)
    order = np.argsort(pts[:, 4], kind="stable")
    ref_sorted, rep = oracle.deskew_compensate(state36, poses, pts[order])
    if first_time > 0:
        assert rep >= 2
    ref = np.zeros_like(ref_sorted)
    ref[order] = ref_sorted
    dm = ScanToMap(lib, max_scan_points=max(n, 1024), max_map_points=4096)
    dm.scan_deskew(pts, poses, state36[:24])
    und = dm.scan_get_undistorted(n)
    # fp64 math, fp32 store: libm vs CUDA sin/cos may differ in the last ulp of the double -> 1 float ulp
    np.testing.assert_allclose(und[:, :3], ref[:, :3], rtol=2e-7, atol=2e-6)
    np.testing.assert_array_equal(und[:, 3], pts[:, 8])
    moved = np.abs(ref[:, :3] - pts[:, :3]).max(1) > 0
    assert moved.sum() > 0.9 * n
    dm.close()
