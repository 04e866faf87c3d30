"""Independent checks of the oracle's restated pieces that the reference itself cannot pin here (Eigen and
PCL are not installed, the reference ships no fixtures for them).  Each restatement is held against a DIFFERENT
implementation of the same published algorithm (LAPACK through numpy / scipy, plain numpy integer arithmetic) and against
the properties the reference's call sites rely on.  CPU only."""
import numpy as np
import pytest
import scipy.linalg
from scipy.spatial.transform import Rotation


def _patches(n, seed, noise=0.01):
    """n near-planar 5-point neighbourhoods the way the map produces them: points of one surface within ~1 m"""
    r = np.random.default_rng(seed)
    nrm = r.normal(size=(n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    centre = r.uniform(-60, 60, (n, 3))
    a = np.cross(nrm, r.normal(size=(n, 3)))
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    b = np.cross(nrm, a)
    uv = r.uniform(-0.8, 0.8, (n, 5, 2))
    pts = centre[:, None, :] + uv[..., :1] * a[:, None, :] + uv[..., 1:] * b[:, None, :] + r.normal(0, noise, (n, 5, 1)) * nrm[:, None, :]
    return pts.astype(np.float32)


def test_esti_plane_against_lapack(oracle):
    """common_lib.h:267-299 solves A n = -1 (A 5x3, float) with Eigen's column-pivoted Householder QR.  LAPACK's sgeqp3 is the
    same Businger-Golub algorithm with its own operation order: the two cannot agree bit for bit, but they must sit at the
    same distance from the fp64 least-squares plane (the problem is ill-conditioned in (n, d) separately -- |centre| / patch
    size ~ 100 -- and well-conditioned in what the residual uses: the plane's distance at the patch)."""
    pts = _patches(1000, 3)
    pabcd, ok = oracle.esti_plane(pts, 0.1)
    assert ok.mean() > 0.98  # (a few random patches are nearly collinear: the noise then tilts the plane past the 0.1 m gate)
    np.testing.assert_allclose(np.linalg.norm(pabcd[:, :3], axis=1), 1.0, atol=3e-7)
    e_n, l_n, e_p, l_p = [], [], [], []
    for i in np.nonzero(ok)[0]:
        A = pts[i]
        rhs = -np.ones(5, np.float32)
        q, rr, piv = scipy.linalg.qr(A, mode="economic", pivoting=True)  # float32 in, float32 LAPACK (sgeqp3)
        assert q.dtype == np.float32
        x = np.zeros(3, np.float32)
        x[piv] = scipy.linalg.solve_triangular(rr, q.T @ rhs)
        nn = np.linalg.norm(x)
        lap = np.concatenate([x / nn, [1.0 / nn]]).astype(np.float64)
        x64 = np.linalg.lstsq(A.astype(np.float64), -np.ones(5), rcond=None)[0]
        n64 = np.linalg.norm(x64)
        ref = np.concatenate([x64 / n64, [1.0 / n64]])
        c = A.astype(np.float64).mean(axis=0)
        mine = pabcd[i].astype(np.float64)
        e_n.append(np.abs(mine[:3] - ref[:3]).max())
        l_n.append(np.abs(lap[:3] - ref[:3]).max())
        e_p.append(abs((mine[:3] @ c + mine[3]) - (ref[:3] @ c + ref[3])))
        l_p.append(abs((lap[:3] @ c + lap[3]) - (ref[:3] @ c + ref[3])))
    e_n, l_n, e_p, l_p = map(np.array, (e_n, l_n, e_p, l_p))
    assert e_p.max() < 1e-4  # the plane itself: within 0.1 mm of the fp64 plane at the patch (measured 2e-5; LAPACK float the same)
    assert np.median(e_n) < 2 * np.median(l_n) + 1e-7 and np.percentile(e_n, 99) < 2 * np.percentile(l_n, 99) + 1e-6
    assert np.median(e_p) < 2 * np.median(l_p) + 1e-7 and np.percentile(e_p, 99) < 2 * np.percentile(l_p, 99) + 1e-6


def test_esti_plane_gate_and_degenerate_inputs(oracle):
    """the validity flag is exactly `all |n . p_j + d| <= threshold` in float (:289-296); rank-deficient neighbourhoods
    (collinear, repeated points) give a finite answer or a rejected plane, never NaN garbage that passes the gate"""
    pts = _patches(600, 5, noise=0.06)  # noisy enough that a good share of patches fails the 0.1 gate
    pabcd, ok = oracle.esti_plane(pts, 0.1)
    assert 0.05 < ok.mean() < 0.999
    res = np.abs((pts.astype(np.float64) * pabcd[:, None, :3]).sum(-1) + pabcd[:, None, 3])
    clear = np.abs(res - 0.1).min(axis=1) > 1e-4  # away from the threshold the float and fp64 evaluations must agree
    np.testing.assert_array_equal(ok[clear], (res[clear] <= 0.1).all(axis=1))
    # a tighter threshold can only reject more
    _, ok2 = oracle.esti_plane(pts, 0.02)
    assert not (ok2 & ~ok).any()
    line = np.zeros((1, 5, 3), np.float32)
    line[0, :, 0] = np.arange(5) + 3.0
    line[0, :, 1] = 2.0 * (np.arange(5) + 3.0)
    same = np.tile(np.array([[1.5, -2.0, 0.7]], np.float32), (1, 5, 1))
    zero = np.zeros((1, 5, 3), np.float32)
    for bad in (line, same):
        p, okb = oracle.esti_plane(bad, 0.1)
        assert np.isfinite(p).all()  # Eigen's rank-revealing solve zeroes the dependent components
        if okb[0]:
            r = np.abs((bad[0].astype(np.float64) * p[0, :3]).sum(-1) + p[0, 3])
            assert (r <= 0.1 + 1e-5).all()
    # reference quirk, reproduced: five points at the origin give rank 0, normvec = 0, pca_result = 0/0 -- and the gate
    # `fabs(NaN) > threshold` (common_lib.h:291) is false, so the reference returns TRUE with a NaN plane.  It stays harmless:
    # pd2 = NaN makes `s > 0.9` (laserMapping.cpp:870) false, the point is not selected.
    p, okb = oracle.esti_plane(zero, 0.1)
    assert okb[0] and np.isnan(p).all()


def test_so3_exp_log(oracle):
    """so3_math.h:54-81: Rodrigues with the 1e-5 angle switch, Log with its trace / small-angle switches"""
    r = np.random.default_rng(11)
    for _ in range(200):
        w = r.normal(size=3) * r.choice([1e-7, 1e-3, 0.3, 2.5])
        R = oracle.so3_exp3(w)
        if np.linalg.norm(w) > 1e-5:
            np.testing.assert_allclose(R, Rotation.from_rotvec(w).as_matrix(), atol=1e-12)
            if np.linalg.norm(w) < 3.0:
                np.testing.assert_allclose(oracle.so3_log(R), w, atol=1e-7 * max(1.0, np.linalg.norm(w)))
        else:
            np.testing.assert_array_equal(R, np.eye(3))  # the reference returns the identity below the switch, not I + [w]x
        np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-12)
    # Exp(w, dt) = Exp(w * dt) (IMU_Processing.hpp:354)
    w = np.array([0.2, -0.1, 0.4])
    np.testing.assert_allclose(oracle.so3_exp(w, 0.05), oracle.so3_exp3(w * 0.05), atol=1e-15)
    np.testing.assert_array_equal(oracle.so3_log(np.eye(3)), np.zeros(3))


def test_jacobi6_against_lapack(oracle):
    """the degeneracy check (new): cyclic Jacobi on the 6x6 pose block, ascending eigenvalues, eigenvectors in columns"""
    r = np.random.default_rng(13)
    for trial in range(50):
        J = r.normal(size=(200, 6)) * np.array([1, 1, 1, 10, 10, 10.0])
        if trial % 5 == 0:
            J[:, 3] = 0.0  # a corridor: no constraint along x
        A = J.T @ J
        ev, V = oracle.jacobi6(A)
        ref = np.linalg.eigvalsh(A)
        assert (np.diff(ev) >= 0).all()
        np.testing.assert_allclose(ev, ref, rtol=1e-9, atol=1e-9 * ref[-1])
        np.testing.assert_allclose(V.T @ V, np.eye(6), atol=1e-10)
        np.testing.assert_allclose(V @ np.diag(ev) @ V.T, A, atol=1e-9 * ref[-1])
        np.testing.assert_allclose(ev.sum(), np.trace(A), rtol=1e-12)
        if trial % 5 == 0:
            assert abs(ev[0]) < 1e-9 * ref[-1]
            assert abs(abs(V[3, 0]) - 1.0) < 1e-9  # the weakest direction is the corridor axis


def _voxel_grid_numpy(xyz, leaf):
    """pcl::VoxelGrid::applyFilter (PCL 1.10 voxel_grid.hpp) in plain numpy: float inverse leaf, floor of the float product,
    int32 index arithmetic, ascending index order, centroid per voxel"""
    inv = np.float32(1.0) / np.float32(leaf)
    mn = xyz.min(axis=0)
    mx = xyz.max(axis=0)
    min_b = np.floor(mn * inv).astype(np.int64)
    max_b = np.floor(mx * inv).astype(np.int64)
    div = max_b - min_b + 1
    ijk = np.floor(xyz * inv).astype(np.int64) - min_b
    idx = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    uniq, inverse = np.unique(idx, return_inverse=True)
    cent = np.zeros((len(uniq), 3))
    np.add.at(cent, inverse, xyz.astype(np.float64))
    cnt = np.bincount(inverse, minlength=len(uniq))
    return cent / cnt[:, None], inverse, cnt


@pytest.mark.parametrize("seed,span,shift", [(1, 8.0, 0.0), (2, 40.0, -300.0), (3, 0.3, 1000.0)])
def test_voxel_grid_restatement_against_numpy(oracle, seed, span, shift):
    r = np.random.default_rng(seed)
    n = 6000
    p = np.zeros((n, 12), np.float32)
    p[:, :3] = r.uniform(-span, span, (n, 3)) + shift
    p[: n // 3, :3] = p[0, :3] + r.uniform(-0.2, 0.2, (n // 3, 3))  # a clump: many points per voxel
    p[:, 8] = r.uniform(0, 255, n)
    out, vop, _ = oracle.voxel_grid(p, 0.5)
    cent, inverse, cnt = _voxel_grid_numpy(p[:, :3], 0.5)
    assert len(out) == len(cent)
    np.testing.assert_array_equal(vop, inverse)  # voxel of every point = rank of its index in ascending order: integer-exact
    np.testing.assert_allclose(out[:, :3], cent, rtol=2e-6, atol=2e-5 + 1e-6 * abs(shift))
    inten = np.bincount(inverse, weights=p[:, 8].astype(np.float64)) / cnt
    np.testing.assert_allclose(out[:, 8], inten, rtol=1e-5, atol=1e-3)  # downsample_all_data_: every field is averaged
    # every centroid lies in the voxel whose points produced it
    inv = np.float32(2.0)
    min_b = np.floor(p[:, :3].min(axis=0) * inv).astype(np.int64)
    member = np.zeros(len(cent), np.int64)
    member[inverse] = np.arange(n)  # any one point of each voxel
    ijk = np.floor(p[member, :3] * inv).astype(np.int64) - min_b
    rel = out[:, :3].astype(np.float64) * 2.0 - min_b
    slack = 1e-3 + 1e-6 * abs(shift)
    assert (rel >= ijk - slack).all() and (rel <= ijk + 1 + slack).all()


def test_voxel_grid_restatement_edge_cases(oracle):
    out, vop, _ = oracle.voxel_grid(np.zeros((0, 12), np.float32), 0.5)
    assert len(out) == 0
    one = np.zeros((1, 12), np.float32)
    one[0, :3] = [1.2, -3.4, 5.6]
    out, vop, _ = oracle.voxel_grid(one, 0.5)
    assert len(out) == 1 and vop[0] == 0
    np.testing.assert_array_equal(out[0, :3], one[0, :3])
    # the int32 overflow guard of applyFilter: (dx * dy * dz) > INT32_MAX -> the cloud passes through unfiltered
    far = np.zeros((4, 12), np.float32)
    far[:, :3] = [[0, 0, 0], [1e5, 0, 0], [0, 1e5, 0], [0, 0, 1e5]]
    out, vop, _ = oracle.voxel_grid(far, 0.5)
    assert len(out) == 4
    np.testing.assert_array_equal(out[:, :3], far[:, :3])
