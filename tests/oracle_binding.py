"""ctypes access to the CPU parity oracle (oracle/liboracle.so, oracle/_ref/libikd_ref.so).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")

MAP_PORT, MAP_REF = 0, 1


class OrcLioConfig(C.Structure):
    _fields_ = [
        ("max_iteration", C.c_int),
        ("filter_size_surf", C.c_double),
        ("filter_size_map", C.c_double),
        ("cube_len", C.c_double),
        ("extrinsic_est_en", C.c_int),
        ("featptsThreshold", C.c_int),
        ("beta", C.c_double),
        ("det_range", C.c_float),
        ("extrinT", C.c_double * 3),
        ("extrinR", C.c_double * 9),
    ]


class OrcThermal(C.Structure):
    _fields_ = [
        ("tis_online", C.c_int),
        ("recv_n", C.c_int),
        ("delta_pos", C.c_double * 3),
        ("delta_quat", C.c_double * 4),
        ("delta_vel", C.c_double * 3),
        ("cov_slots", C.c_double * 8),
        ("l2l_pos", C.c_double * 3),
        ("l2l_quat", C.c_double * 4),
        ("l2l_vel", C.c_double * 3),
        ("l2l_cov_slots", C.c_double * 8),
    ]


class OrcScanSummary(C.Structure):
    _fields_ = [
        ("had_points", C.c_int), ("built_map", C.c_int), ("did_update", C.c_int), ("ekf_stop", C.c_int),
        ("n_raw", C.c_int), ("n_down", C.c_int), ("map_points_before", C.c_int), ("deleted", C.c_int),
        ("added", C.c_int), ("n_iters", C.c_int), ("n_added_ds", C.c_int), ("n_added_raw", C.c_int),
        ("eigvals", C.c_double * 6), ("eigvecs", C.c_double * 36), ("state_prop", C.c_double * 36),
        ("t_deskew", C.c_double), ("t_voxel", C.c_double), ("t_knn", C.c_double), ("t_resid", C.c_double),
        ("t_solve", C.c_double), ("t_insert", C.c_double), ("t_delete", C.c_double),
    ]


class OrcIter(C.Structure):
    _fields_ = [
        ("iter", C.c_int), ("effct_feat_num", C.c_int), ("converged", C.c_int), ("ekf_stop", C.c_int),
        ("did_match", C.c_int), ("n_down", C.c_int),
        ("total_residual", C.c_double), ("res_mean_last", C.c_double),
        ("HtH", C.c_double * 144), ("Htr", C.c_double * 12), ("pose_in", C.c_double * 24),
        ("state_out", C.c_double * 36), ("solution", C.c_double * 24),
    ]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def default_lio_config(**kw) -> OrcLioConfig:
    c = OrcLioConfig()
    c.max_iteration = 4
    c.filter_size_surf = 0.5
    c.filter_size_map = 0.5
    c.cube_len = 1000.0
    c.extrinsic_est_en = 0
    c.featptsThreshold = 30
    c.beta = 0.1
    c.det_range = 300.0
    for i in range(3):
        c.extrinT[i] = 0.0
    for i, v in enumerate([1, 0, 0, 0, 1, 0, 0, 0, 1]):
        c.extrinR[i] = float(v)
    for k, v in kw.items():
        if k in ("extrinT", "extrinR"):
            for i, x in enumerate(v):
                getattr(c, k)[i] = float(x)
        else:
            setattr(c, k, v)
    return c


class OracleMap:
    def __init__(self, orc, kind: int, ds: float, handle=None, owned=True):
        self.o = orc
        self.owned = owned
        self.h = handle if handle is not None else orc.lib.orc_map_create(C.c_int(kind), C.c_float(ds))
        if not self.h:
            raise RuntimeError("oracle map backend unavailable (reference library not built?)")

    def close(self):
        if self.h and self.owned:
            self.o.lib.orc_map_destroy(C.c_void_p(self.h))
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def build(self, xyzi):
        a = np.ascontiguousarray(xyzi, np.float32).reshape(-1, 4)
        self.o.lib.orc_map_build(C.c_void_p(self.h), _p(a), C.c_int(len(a)))

    def knn(self, q, k=5):
        q = np.ascontiguousarray(q, np.float32).reshape(-1, 3)
        n = len(q)
        pts = np.zeros((n, k, 4), np.float32)
        d2 = np.zeros((n, k), np.float32)
        cnt = np.zeros(n, np.int32)
        self.o.lib.orc_map_knn(C.c_void_p(self.h), _p(q), C.c_int(n), C.c_int(k), _p(pts), _p(d2), _p(cnt))
        return pts, d2, cnt

    def add(self, xyzi, downsample: bool) -> int:
        a = np.ascontiguousarray(xyzi, np.float32).reshape(-1, 4)
        if len(a) == 0:
            return 0
        return self.o.lib.orc_map_add(C.c_void_p(self.h), _p(a), C.c_int(len(a)), C.c_int(1 if downsample else 0))

    def delete_boxes(self, boxes) -> int:
        b = np.ascontiguousarray(boxes, np.float32).reshape(-1, 6)
        return self.o.lib.orc_map_delete_boxes(C.c_void_p(self.h), _p(b), C.c_int(len(b)))

    def _settle(self):
        """The reference ikd-Tree rebuilds unbalanced sub-trees on a background pthread (ikd_Tree.cpp:229-367) and
        its flatten / validnum are not atomic against it (the reference itself races here, laserMapping.cpp:1172).
        Give that thread a moment before reading whole-tree contents so set comparisons are deterministic."""
        import time

        time.sleep(0.05)

    def validnum(self) -> int:
        self._settle()
        return self.o.lib.orc_map_validnum(C.c_void_p(self.h))

    def flatten(self) -> np.ndarray:
        self._settle()
        n = self.o.lib.orc_map_flatten(C.c_void_p(self.h), None, C.c_int(0))
        out = np.zeros((max(n, 1), 4), np.float32)
        m = self.o.lib.orc_map_flatten(C.c_void_p(self.h), _p(out), C.c_int(len(out)))
        return out[: min(n, m)]


class OracleLio:
    """The restated per-scan update (oracle/oracle.cpp)."""

    def __init__(self, orc, cfg: OrcLioConfig | None = None, map_kind: int = MAP_REF):
        self.o = orc
        self.cfg = cfg or default_lio_config()
        self.h = orc.lib.orc_lio_create(C.byref(self.cfg), C.c_int(map_kind))
        if not self.h:
            raise RuntimeError("oracle lio unavailable")
        self.summary = OrcScanSummary()

    def close(self):
        if self.h:
            self.o.lib.orc_lio_destroy(C.c_void_p(self.h))
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_threads(self, n):
        self.o.lib.orc_lio_set_threads(C.c_void_p(self.h), C.c_int(n))

    def on_lidar_msg(self):
        self.o.lib.orc_lio_on_lidar_msg(C.c_void_p(self.h))

    def on_edge_count(self, n):
        self.o.lib.orc_lio_on_edge_count(C.c_void_p(self.h), C.c_int(n))

    def force_imu_ready(self, mean_acc, last_imu7):
        a = np.ascontiguousarray(mean_acc, np.float64)
        b = np.ascontiguousarray(last_imu7, np.float64)
        self.o.lib.orc_lio_force_imu_ready(C.c_void_p(self.h), _p(a), _p(b))

    def get_state(self) -> np.ndarray:
        s = np.zeros(36 + 576, np.float64)
        self.o.lib.orc_lio_get_state(C.c_void_p(self.h), _p(s))
        return s

    def set_state(self, s, also_last=True):
        s = np.ascontiguousarray(s, np.float64)
        assert s.size == 612
        self.o.lib.orc_lio_set_state(C.c_void_p(self.h), _p(s))
        if also_last:
            self.o.lib.orc_lio_set_last_state(C.c_void_p(self.h), _p(s))

    def map(self) -> OracleMap:
        return OracleMap(self.o, 0, 0.0, handle=self.o.lib.orc_lio_map(C.c_void_p(self.h)), owned=False)

    def process_scan(self, pts48, lidar_beg_time, imu7, thermal: OrcThermal | None = None) -> OrcScanSummary:
        a = np.ascontiguousarray(pts48, np.float32).reshape(-1, 12)
        im = np.ascontiguousarray(imu7, np.float64).reshape(-1, 7)
        th = C.byref(thermal) if thermal is not None else None
        self.o.lib.orc_lio_process_scan(C.c_void_p(self.h), _p(a), C.c_int(len(a)), C.c_double(lidar_beg_time), _p(im),
                                        C.c_int(len(im)), th, C.byref(self.summary))
        return self.summary

    def iters(self):
        n = self.summary.n_iters
        arr = (OrcIter * max(n, 1))()
        self.o.lib.orc_lio_get_iters(C.c_void_p(self.h), arr, C.c_int(n))
        return [arr[i] for i in range(n)]

    def undistort(self) -> np.ndarray:
        n = self.o.lib.orc_lio_get_undistort(C.c_void_p(self.h), None, C.c_int(0))
        out = np.zeros((max(n, 1), 12), np.float32)
        self.o.lib.orc_lio_get_undistort(C.c_void_p(self.h), _p(out), C.c_int(len(out)))
        return out[:n]

    def feats_down(self) -> np.ndarray:
        n = self.o.lib.orc_lio_get_feats_down(C.c_void_p(self.h), None, C.c_int(0))
        out = np.zeros((max(n, 1), 12), np.float32)
        self.o.lib.orc_lio_get_feats_down(C.c_void_p(self.h), _p(out), C.c_int(len(out)))
        return out[:n]

    def nearest(self):
        n = self.summary.n_down
        near = np.zeros((max(n, 1), 5, 4), np.float32)
        d2 = np.zeros((max(n, 1), 5), np.float32)
        cnt = np.zeros(max(n, 1), np.int32)
        sel = np.zeros(max(n, 1), np.uint8)
        m = self.o.lib.orc_lio_get_nearest(C.c_void_p(self.h), _p(near), _p(d2), _p(cnt), _p(sel), C.c_int(n))
        m = min(m, n)
        return near[:m], d2[:m], cnt[:m], sel[:m]

    def added(self):
        nd, nr = self.summary.n_added_ds, self.summary.n_added_raw
        a = np.zeros((max(nd, 1), 4), np.float32)
        b = np.zeros((max(nr, 1), 4), np.float32)
        self.o.lib.orc_lio_get_added(C.c_void_p(self.h), _p(a), C.c_int(nd), _p(b), C.c_int(nr))
        return a[:nd], b[:nr]

    def imu_poses(self) -> np.ndarray:
        out = np.zeros((512, 22), np.float64)
        n = self.o.lib.orc_lio_get_imu_poses(C.c_void_p(self.h), _p(out), C.c_int(512))
        return out[:n]

    def flags(self):
        f = np.zeros(8, np.int32)
        self.o.lib.orc_lio_get_flags(C.c_void_p(self.h), _p(f))
        return dict(ekf_stop=int(f[0]), ekf_inited=int(f[1]), threshold=int(f[2]), lidar_cnt=int(f[3]),
                    localmap_init=int(f[4]), queue=int(f[5]), first_point_repeat=int(f[6]))

    def localmap(self):
        b = np.zeros(6, np.float32)
        self.o.lib.orc_lio_get_localmap(C.c_void_p(self.h), _p(b))
        return b

    def imu_ready(self) -> bool:
        return bool(self.o.lib.orc_lio_imu_ready(C.c_void_p(self.h)))


class Oracle:
    def __init__(self, lib, ref_ok):
        self.lib = lib
        self.ref_ok = ref_ok

    def esti_plane(self, pts, thr=0.1):
        a = np.ascontiguousarray(pts, np.float32).reshape(-1, 5, 3)
        n = len(a)
        pabcd = np.zeros((n, 4), np.float32)
        ok = np.zeros(n, np.uint8)
        self.lib.orc_esti_plane_batch(_p(a), C.c_int(n), C.c_float(thr), _p(pabcd), _p(ok))
        return pabcd, ok.astype(bool)

    def frontend_sample(self, cloud, layout, sensor, point_filter_num=5, min_range=0.5, max_range=1000.0):
        """cachePointCloud + samplePointCloud restated (oracle.hpp): returns (records (m,12) float32, timespan, stamp_shift)"""
        buf = np.ascontiguousarray(cloud, np.uint8).reshape(-1)
        lay = (C.c_int * 7)(*[int(v) for v in layout])
        n = buf.size // int(layout[0])
        out = np.zeros((max(n, 1), 12), np.float32)
        ts, sh = C.c_double(0), C.c_double(0)
        m = self.lib.orc_frontend_sample(_p(buf), C.c_int(n), lay, C.c_int(sensor), C.c_int(point_filter_num), C.c_float(min_range),
                                         C.c_float(max_range), _p(out), C.c_int(len(out)), C.byref(ts), C.byref(sh))
        return out[:m], ts.value, sh.value

    def voxel_grid(self, pts48, leaf=0.5, stable=False):
        a = np.ascontiguousarray(pts48, np.float32).reshape(-1, 12)
        n = len(a)
        out = np.zeros((max(n, 1), 12), np.float32)
        vop = np.zeros(max(n, 1), np.int32)
        iop = np.zeros(max(n, 1), np.uint32)
        m = self.lib.orc_voxel_grid(_p(a), C.c_int(n), C.c_float(leaf), C.c_int(1 if stable else 0), _p(out), C.c_int(len(out)),
                                    _p(vop), _p(iop))
        return out[:m], vop[:n], iop[:n]

    def deskew_compensate(self, state36, poses22, pts48_sorted):
        st = np.ascontiguousarray(state36, np.float64).reshape(36)
        ps = np.ascontiguousarray(poses22, np.float64).reshape(-1, 22)
        pts = np.ascontiguousarray(pts48_sorted, np.float32).reshape(-1, 12).copy()
        rep = self.lib.orc_deskew_compensate(_p(st), _p(ps), C.c_int(len(ps)), _p(pts), C.c_int(len(pts)))
        return pts, rep

    def so3_exp(self, w, dt):
        w = np.ascontiguousarray(w, np.float64)
        R = np.zeros(9)
        self.lib.orc_so3_exp(_p(w), C.c_double(dt), _p(R))
        return R.reshape(3, 3)

    def so3_exp3(self, w):
        w = np.ascontiguousarray(w, np.float64)
        R = np.zeros(9)
        self.lib.orc_so3_exp3(_p(w), _p(R))
        return R.reshape(3, 3)

    def so3_log(self, R):
        R = np.ascontiguousarray(R, np.float64).reshape(9)
        w = np.zeros(3)
        self.lib.orc_so3_log(_p(R), _p(w))
        return w

    def jacobi6(self, A):
        A = np.ascontiguousarray(A, np.float64).reshape(36)
        ev = np.zeros(6)
        vec = np.zeros(36)
        self.lib.orc_jacobi6(_p(A), _p(ev), _p(vec))
        return ev, vec.reshape(6, 6)

    def inverse(self, A):
        A = np.ascontiguousarray(A, np.float64)
        n = A.shape[0]
        out = np.zeros((n, n))
        ok = self.lib.orc_inverse(_p(A), C.c_int(n), _p(out))
        return out, bool(ok)

    def new_map(self, kind, ds=0.5) -> OracleMap:
        return OracleMap(self, kind, ds)

    def new_lio(self, cfg=None, map_kind=MAP_REF) -> OracleLio:
        return OracleLio(self, cfg, map_kind)


_cached = None


def load(build=True) -> Oracle:
    global _cached
    if _cached is not None:
        return _cached
    so = os.path.join(ODIR, "liboracle.so")
    if build:
        subprocess.run(["make", "-s", "-C", ODIR], check=True, stdout=subprocess.DEVNULL)
    lib = C.CDLL(so)
    lib.orc_map_create.restype = C.c_void_p
    lib.orc_lio_create.restype = C.c_void_p
    lib.orc_lio_map.restype = C.c_void_p
    ref = os.path.join(ODIR, "_ref", "libikd_ref.so")
    ref_ok = os.path.exists(ref) and lib.orc_load_ref(ref.encode()) == 1
    _cached = Oracle(lib, ref_ok)
    return _cached
