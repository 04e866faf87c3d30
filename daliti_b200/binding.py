"""ctypes binding of include/daliti_b200.h (the device-path C ABI).

Mirrors the reference-side variables that cross the cut in
eskf_lio/src/laserMapping.cpp:820-979.  Plumbing only: every
computation happens in the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def default_library_path() -> str:
    return os.path.join(_HERE, "lib", "libdaliti_b200.so")


class DltError(RuntimeError):
    pass


class DltConfig(C.Structure):
    _fields_ = [
        ("ds_scan", C.c_float),
        ("ds_map", C.c_float),
        ("max_sq_dist", C.c_float),
        ("plane_thr", C.c_float),
        ("extrinsic_est_en", C.c_int),
        ("device", C.c_int),
        ("max_scan_points", C.c_int),
        ("max_map_points", C.c_int),
        ("voxel_bitmap_bits", C.c_longlong),
        ("shard_rank", C.c_int),
        ("shard_count", C.c_int),
        ("shard_tile_shift", C.c_int),
    ]


class _MeasureOut(C.Structure):
    _fields_ = [
        ("HtH", C.c_double * 144),
        ("Htr", C.c_double * 12),
        ("total_residual", C.c_double),
        ("effct_feat_num", C.c_int),
        ("n_down", C.c_int),
        ("n_unresolved", C.c_int),
        ("reserved", C.c_int),
    ]


@dataclass
class Measurement:
    HtH: np.ndarray
    Htr: np.ndarray
    total_residual: float
    effct_feat_num: int
    n_down: int
    n_unresolved: int


PEER_BLOB_BYTES = 128  # DLT_PEER_BLOB_BYTES

_ERR = {1: "invalid argument", 2: "no usable CUDA device (there is no CPU path)", 3: "CUDA error", 4: "capacity exceeded", 5: "call-sequence error"}

_SYMBOLS = [
    "dlt_default_config", "dlt_create", "dlt_destroy", "dlt_last_error", "dlt_set_stream", "dlt_sync",
    "dlt_map_build", "dlt_map_build_from_scan", "dlt_map_add", "dlt_map_delete_boxes", "dlt_map_valid_count", "dlt_map_export", "dlt_map_knn",
    "dlt_scan_deskew", "dlt_scan_deskew_dev", "dlt_scan_downsample", "dlt_scan_get_undistorted", "dlt_scan_get_down", "dlt_scan_set_down",
    "dlt_scan_get_voxel_of_point", "dlt_measure", "dlt_measure_dev", "dlt_effective_points", "dlt_get_nearest",
    "dlt_fetch_result", "dlt_degeneracy", "dlt_degeneracy_begin", "dlt_map_incremental", "dlt_set_profiling", "dlt_get_profile", "dlt_launch_count",
    "dlt_scan_downsample_async", "dlt_iekf_update", "dlt_get_timeline", "dlt_get_iekf_clocks", "dlt_set_shard_reduce", "dlt_result_dev", "dlt_frontend_sample", "dlt_frontend_read", "dlt_scan_prefetch",
    "dlt_peer_export", "dlt_peer_attach", "dlt_peer_detach", "dlt_map_incremental_async", "dlt_map_incremental_collect", "dlt_debug_counters",
]


def load_library(path: str | None = None) -> C.CDLL:
    """Load the CUDA library.  Raises DltError when it has not been built."""
    path = path or default_library_path()
    if not os.path.exists(path):
        raise DltError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  daliti_b200 has no CPU fallback."
        )
    lib = C.CDLL(path)
    for s in _SYMBOLS:
        if not hasattr(lib, s):
            raise DltError(f"{path} does not export {s}")
    lib.dlt_last_error.restype = C.c_char_p
    lib.dlt_last_error.argtypes = [C.c_void_p]
    lib.dlt_launch_count.restype = C.c_ulonglong
    return lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class ScanToMap:
    """One handle of the device path: map + current scan + measurement model."""

    def __init__(self, lib: C.CDLL | None = None, **cfg):
        self.lib = lib or load_library()
        c = DltConfig()
        self.lib.dlt_default_config(C.byref(c))
        for k, v in cfg.items():
            if not hasattr(c, k):
                raise TypeError(f"unknown config field {k}")
            setattr(c, k, v)
        self.cfg = c
        self.h = C.c_void_p()
        rc = self.lib.dlt_create(C.byref(c), C.byref(self.h))
        if rc != 0:
            self.h = None
            raise DltError(f"dlt_create failed: {_ERR.get(rc, rc)}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.dlt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            msg = self.lib.dlt_last_error(self.h)
            raise DltError(f"{_ERR.get(rc, rc)}: {msg.decode() if msg else ''}")

    # ---- map
    def map_build(self, xyzi):
        a = _f32(xyzi).reshape(-1, 4)
        self._ck(self.lib.dlt_map_build(self.h, _p(a), C.c_int(a.shape[0])))

    def map_build_from_scan(self, pose24):
        ps = _f64(pose24).reshape(24)
        self._ck(self.lib.dlt_map_build_from_scan(self.h, _p(ps)))

    def map_add(self, xyzi, downsample: bool):
        a = _f32(xyzi).reshape(-1, 4)
        self._ck(self.lib.dlt_map_add(self.h, _p(a), C.c_int(a.shape[0]), C.c_int(1 if downsample else 0)))

    def map_delete_boxes(self, boxes) -> int:
        b = _f32(boxes).reshape(-1, 6)
        d = C.c_int(0)
        self._ck(self.lib.dlt_map_delete_boxes(self.h, _p(b), C.c_int(b.shape[0]), C.byref(d)))
        return d.value

    def map_valid_count(self) -> int:
        n = C.c_int(0)
        self._ck(self.lib.dlt_map_valid_count(self.h, C.byref(n)))
        return n.value

    def map_export(self) -> np.ndarray:
        n = self.map_valid_count()
        out = np.zeros((max(n, 1), 4), np.float32)
        m = C.c_int(0)
        self._ck(self.lib.dlt_map_export(self.h, _p(out), C.c_int(out.shape[0]), C.byref(m)))
        return out[: min(m.value, n)]

    def map_knn(self, q_xyz):
        q = _f32(q_xyz).reshape(-1, 3)
        nq = q.shape[0]
        pts = np.zeros((nq, 5, 4), np.float32)
        d2 = np.zeros((nq, 5), np.float32)
        cnt = np.zeros(nq, np.int32)
        self._ck(self.lib.dlt_map_knn(self.h, _p(q), C.c_int(nq), _p(pts), _p(d2), _p(cnt)))
        return pts, d2, cnt

    # ---- front end (feature_extract.cpp:264-450 with feature_enabled = 0)
    SENSORS = {"velodyne": 0, "livox": 1, "ouster": 2, "robosense": 3}

    def frontend_sample(self, cloud: np.ndarray, layout, sensor, point_filter_num=5, min_range=0.5, max_range=1000.0):
        """cloud: the PointCloud2 data as uint8 (n_points * point_step); layout: (point_step, off_x, off_y, off_z, off_intensity, off_ring,
        off_time).  Returns (device pointer to the 48-byte records, n, timespan, sweep_span, stamp_shift)."""
        buf = np.ascontiguousarray(cloud, dtype=np.uint8).reshape(-1)
        lay = (C.c_int * 7)(*[int(v) for v in layout])
        n_points = buf.size // int(layout[0])
        ptr, n = C.c_void_p(), C.c_int(0)
        ts, sp, sh = C.c_double(0), C.c_double(0), C.c_double(0)
        code = self.SENSORS[sensor] if isinstance(sensor, str) else int(sensor)
        self._ck(self.lib.dlt_frontend_sample(self.h, _p(buf), C.c_int(n_points), lay, C.c_int(code), C.c_int(point_filter_num), C.c_float(min_range),
                                              C.c_float(max_range), C.byref(ptr), C.byref(n), C.byref(ts), C.byref(sp), C.byref(sh)))
        return ptr.value, n.value, ts.value, sp.value, sh.value

    def frontend_read(self, n) -> np.ndarray:
        out = np.zeros((max(n, 1), 12), np.float32)
        self._ck(self.lib.dlt_frontend_read(self.h, _p(out), C.c_int(n)))
        return out[:n]

    # ---- scan
    def scan_deskew(self, pts48, imu_pose22=None, pose24=None):
        a = np.ascontiguousarray(pts48, dtype=np.float32).reshape(-1, 12)
        if imu_pose22 is None:
            self._ck(self.lib.dlt_scan_deskew(self.h, _p(a), C.c_int(a.shape[0]), None, C.c_int(0), None))
        else:
            ip = _f64(imu_pose22).reshape(-1, 22)
            ps = _f64(pose24).reshape(24)
            self._ck(self.lib.dlt_scan_deskew(self.h, _p(a), C.c_int(a.shape[0]), _p(ip), C.c_int(ip.shape[0]), _p(ps)))

    def scan_downsample(self) -> int:
        n = C.c_int(0)
        self._ck(self.lib.dlt_scan_downsample(self.h, C.byref(n)))
        return n.value

    def scan_get_undistorted(self, n_raw) -> np.ndarray:
        out = np.zeros((max(n_raw, 1), 4), np.float32)
        n = C.c_int(0)
        self._ck(self.lib.dlt_scan_get_undistorted(self.h, _p(out), C.c_int(out.shape[0]), C.byref(n)))
        return out[: n.value]

    def scan_get_down(self, cap) -> np.ndarray:
        out = np.zeros((max(cap, 1), 4), np.float32)
        n = C.c_int(0)
        self._ck(self.lib.dlt_scan_get_down(self.h, _p(out), C.c_int(out.shape[0]), C.byref(n)))
        return out[: n.value]

    def scan_set_down(self, xyzi):
        a = _f32(xyzi).reshape(-1, 4)
        self._ck(self.lib.dlt_scan_set_down(self.h, _p(a), C.c_int(a.shape[0])))

    def scan_get_voxel_of_point(self, n_raw) -> np.ndarray:
        out = np.zeros(max(n_raw, 1), np.int32)
        self._ck(self.lib.dlt_scan_get_voxel_of_point(self.h, _p(out), C.c_int(n_raw)))
        return out[:n_raw]

    # ---- measurement model
    def measure(self, pose24, do_match: bool) -> Measurement:
        ps = _f64(pose24).reshape(24)
        o = _MeasureOut()
        self._ck(self.lib.dlt_measure(self.h, _p(ps), C.c_int(1 if do_match else 0), C.byref(o)))
        return Measurement(
            HtH=np.array(o.HtH, dtype=np.float64).reshape(12, 12),
            Htr=np.array(o.Htr, dtype=np.float64),
            total_residual=o.total_residual,
            effct_feat_num=o.effct_feat_num,
            n_down=o.n_down,
            n_unresolved=o.n_unresolved,
        )

    def measure_dev(self, pose24, do_match: bool, result_dev_ptr: int):
        ps = _f64(pose24).reshape(24)
        self._ck(self.lib.dlt_measure_dev(self.h, _p(ps), C.c_int(1 if do_match else 0), C.c_void_p(result_dev_ptr)))

    def effective_points(self, cap):
        xyzi = np.zeros((max(cap, 1), 4), np.float32)
        coeff = np.zeros((max(cap, 1), 4), np.float32)
        n = C.c_int(0)
        self._ck(self.lib.dlt_effective_points(self.h, _p(xyzi), _p(coeff), C.c_int(cap), C.byref(n)))
        return xyzi[: n.value], coeff[: n.value]

    def get_nearest(self, n_down):
        nbr = np.zeros((max(n_down, 1), 5, 4), np.float32)
        cnt = np.zeros(max(n_down, 1), np.int32)
        sel = np.zeros(max(n_down, 1), np.uint8)
        self._ck(self.lib.dlt_get_nearest(self.h, _p(nbr), _p(cnt), _p(sel), C.c_int(n_down)))
        return nbr[:n_down], cnt[:n_down], sel[:n_down]

    def degeneracy(self):
        ev = np.zeros(6, np.float64)
        vec = np.zeros(36, np.float64)
        self._ck(self.lib.dlt_degeneracy(self.h, _p(ev), _p(vec)))
        return ev, vec.reshape(6, 6)

    def debug_counters(self) -> np.ndarray:
        out = np.zeros(16, np.int32)
        self._ck(self.lib.dlt_debug_counters(self.h, _p(out)))
        return out

    def map_incremental(self, pose24, flg_EKF_inited: bool = True):
        ps = _f64(pose24).reshape(24)
        a, b = C.c_int(0), C.c_int(0)
        self._ck(self.lib.dlt_map_incremental(self.h, _p(ps), C.c_int(1 if flg_EKF_inited else 0), C.byref(a), C.byref(b)))
        return a.value, b.value

    PROFILE_GROUPS = ("knn", "residual", "deskew", "voxelgrid", "insert", "far_fallback", "iekf_step", "knn8")

    # ---- sharded map over NVLink peer memory (dlt_peer_*): see daliti_b200.sharded.attach_peers
    def peer_export(self) -> bytes:
        blob = (C.c_ubyte * PEER_BLOB_BYTES)()
        self._ck(self.lib.dlt_peer_export(self.h, blob))
        return bytes(blob)

    def peer_attach(self, blobs):
        """blobs: the peer_export() results of all shard_count ranks, in rank order"""
        raw = b"".join(blobs)
        buf = (C.c_ubyte * len(raw)).from_buffer_copy(raw)
        self._ck(self.lib.dlt_peer_attach(self.h, buf))

    def peer_detach(self):
        self._ck(self.lib.dlt_peer_detach(self.h))

    def set_profiling(self, on: bool):
        self._ck(self.lib.dlt_set_profiling(self.h, C.c_int(1 if on else 0)))

    def get_profile(self, reset=True) -> dict:
        ms = np.zeros(8, np.float64)
        cnt = np.zeros(8, np.int64)
        self._ck(self.lib.dlt_get_profile(self.h, _p(ms), _p(cnt), C.c_int(1 if reset else 0)))
        return {g: (float(ms[i]), int(cnt[i])) for i, g in enumerate(self.PROFILE_GROUPS)}

    def get_timeline(self, cap=4096):
        buf = np.zeros((cap, 3), np.float64)
        n = C.c_int(0)
        self._ck(self.lib.dlt_get_timeline(self.h, _p(buf), C.c_int(cap), C.byref(n)))
        names = self.PROFILE_GROUPS + ("eigen6",)
        return [(names[int(k)], float(a), float(b)) for k, a, b in buf[: min(n.value, cap)]]

    def launch_count(self) -> int:
        return int(self.lib.dlt_launch_count())

    def set_stream(self, cuda_stream_ptr: int | None):
        self._ck(self.lib.dlt_set_stream(self.h, C.c_void_p(cuda_stream_ptr) if cuda_stream_ptr else None))

    def sync(self):
        self._ck(self.lib.dlt_sync(self.h))
