"""ctypes binding of include/daliti_b200_lio.h: the host-side per-scan update
(eskf_lio/src/laserMapping.cpp:731-1177 mirrored in C++ over the device C ABI)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .binding import DltConfig, DltError, ScanToMap, _ERR, _f64, _p, load_library


class LioConfig(C.Structure):
    _fields_ = [
        ("dev", DltConfig),
        ("max_iteration", C.c_int),
        ("cube_len", C.c_double),
        ("featptsThreshold", C.c_int),
        ("beta", C.c_double),
        ("det_range", C.c_float),
        ("extrinT", C.c_double * 3),
        ("extrinR", C.c_double * 9),
        ("degeneracy_eig_threshold", C.c_double),
        ("device_loop", C.c_int),
        ("async_insert", C.c_int),
    ]


class LioThermal(C.Structure):
    _fields_ = [
        ("tis_online", C.c_int), ("recv_n", C.c_int),
        ("delta_pos", C.c_double * 3), ("delta_quat", C.c_double * 4), ("delta_vel", C.c_double * 3), ("cov_slots", C.c_double * 8),
        ("l2l_pos", C.c_double * 3), ("l2l_quat", C.c_double * 4), ("l2l_vel", C.c_double * 3), ("l2l_cov_slots", C.c_double * 8),
    ]


class LioIter(C.Structure):
    _fields_ = [
        ("iter", C.c_int), ("effct_feat_num", C.c_int), ("converged", C.c_int), ("ekf_stop", C.c_int),
        ("did_match", C.c_int), ("n_down", C.c_int),
        ("total_residual", C.c_double), ("res_mean_last", C.c_double),
        ("HtH", C.c_double * 144), ("Htr", C.c_double * 12), ("pose_in", C.c_double * 24),
        ("state_out", C.c_double * 36), ("solution", C.c_double * 24),
    ]


class LioScanOut(C.Structure):
    _fields_ = [
        ("had_points", C.c_int), ("built_map", C.c_int), ("did_update", C.c_int), ("ekf_stop", C.c_int),
        ("n_raw", C.c_int), ("n_down", C.c_int), ("map_points_before", C.c_int), ("deleted", C.c_int),
        ("added", C.c_int), ("n_iters", C.c_int), ("n_added_ds", C.c_int), ("n_added_raw", C.c_int),
        ("eigvals", C.c_double * 6), ("eigvecs", C.c_double * 36), ("state_prop", C.c_double * 36),
        ("degenerate", C.c_int), ("reserved", C.c_int),
        ("t_deskew", C.c_double), ("t_voxel", C.c_double), ("t_iterate", C.c_double), ("t_insert", C.c_double),
        ("t_delete", C.c_double), ("t_total", C.c_double),
    ]


_LIO_SYMBOLS = [
    "dlt_lio_default_config", "dlt_lio_create", "dlt_lio_destroy", "dlt_lio_last_error", "dlt_lio_device", "dlt_lio_on_lidar_msg",
    "dlt_lio_on_edge_count", "dlt_lio_force_imu_ready", "dlt_lio_get_state", "dlt_lio_set_state", "dlt_lio_get_flags",
    "dlt_lio_get_localmap", "dlt_lio_set_reduce", "dlt_lio_process_scan", "dlt_lio_process_scan_dev", "dlt_lio_process_cloud", "dlt_lio_prefetch_scan", "dlt_lio_get_iters", "dlt_lio_get_imu_poses",
    "dlt_lio_peer_export", "dlt_lio_peer_attach", "dlt_lio_peer_detach", "dlt_lio_collect_insert", "dlt_lio_replay_sequences",
]


class LioSeqScan(C.Structure):
    _fields_ = [("pts48", C.c_void_p), ("n", C.c_int), ("pts_on_device", C.c_int), ("lidar_beg_time", C.c_double),
                ("observation_end_time", C.c_double), ("imu7", C.c_void_p), ("n_imu", C.c_int), ("lidar_msgs", C.c_int)]


class LioSeq(C.Structure):
    _fields_ = [("h", C.c_void_p), ("scans", C.POINTER(LioSeqScan)), ("n_scans", C.c_int), ("prefetch", C.c_int),
                ("thermal", C.c_void_p), ("outs", C.POINTER(LioScanOut)), ("rc", C.c_int), ("n_done", C.c_int)]


def replay_sequences(lms, scan_lists, n_threads=0, prefetch=False, want_outs=True):
    """dlt_lio_replay_sequences: lms[i] replays scan_lists[i] = [(pts48, t_beg, imu7[, observation_end_time]), ...] natively, all
    sequences concurrently (pts48: numpy / pinned torch CPU tensor, or a CUDA tensor).  Returns per-sequence lists of LioScanOut."""
    lib = lms[0].lib
    keep, seqs = [], (LioSeq * len(lms))()
    outs_all = []
    for i, (lm, scans) in enumerate(zip(lms, scan_lists)):
        arr = (LioSeqScan * max(1, len(scans)))()
        for k, sc in enumerate(scans):
            pts, t_beg, imu = sc[0], sc[1], sc[2]
            on_dev = bool(getattr(pts, "is_cuda", False))
            if hasattr(pts, "data_ptr"):
                n, ptr = int(pts.shape[0]), pts.data_ptr()
            else:
                a = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 12)
                keep.append(a)
                n, ptr = a.shape[0], a.ctypes.data
            im = _f64(imu).reshape(-1, 7)
            keep.append(im)
            arr[k].pts48, arr[k].n, arr[k].pts_on_device = ptr, n, 1 if on_dev else 0
            arr[k].lidar_beg_time = float(t_beg)
            arr[k].observation_end_time = float(sc[3]) if len(sc) > 3 else float(t_beg)
            arr[k].imu7, arr[k].n_imu, arr[k].lidar_msgs = im.ctypes.data, im.shape[0], 1
        outs = (LioScanOut * max(1, len(scans)))()
        keep += [arr, outs]
        outs_all.append(outs)
        seqs[i].h, seqs[i].scans, seqs[i].n_scans = lm.h, arr, len(scans)
        seqs[i].prefetch, seqs[i].thermal = 1 if prefetch else 0, None
        seqs[i].outs = outs if want_outs else None
    rc = lib.dlt_lio_replay_sequences(seqs, C.c_int(len(lms)), C.c_int(n_threads))
    if rc != 0:
        bad = [i for i in range(len(lms)) if seqs[i].rc != 0]
        msg = lib.dlt_lio_last_error(lms[bad[0]].h) if bad else b""
        raise DltError(f"dlt_lio_replay_sequences: {_ERR.get(rc, rc)} (sequences {bad}): {msg.decode() if msg else ''}")
    return [[o[k] for k in range(len(sl))] for o, sl in zip(outs_all, scan_lists)]


class LaserMapping:
    """The reference-facing call: one `process_scan` per LiDAR scan."""

    def __init__(self, lib: C.CDLL | None = None, dev: dict | None = None, **cfg):
        self.lib = lib or load_library()
        for s in _LIO_SYMBOLS:
            if not hasattr(self.lib, s):
                raise DltError(f"library does not export {s}")
        self.lib.dlt_lio_last_error.restype = C.c_char_p
        self.lib.dlt_lio_last_error.argtypes = [C.c_void_p]
        self.lib.dlt_lio_device.restype = C.c_void_p
        self.lib.dlt_lio_device.argtypes = [C.c_void_p]
        c = LioConfig()
        self.lib.dlt_lio_default_config(C.byref(c))
        for k, v in (dev or {}).items():
            if not hasattr(c.dev, k):
                raise TypeError(f"unknown device config field {k}")
            setattr(c.dev, k, v)
        for k, v in cfg.items():
            if k in ("extrinT", "extrinR"):
                for i, x in enumerate(v):
                    getattr(c, k)[i] = float(x)
            elif hasattr(c, k):
                setattr(c, k, v)
            else:
                raise TypeError(f"unknown config field {k}")
        self.cfg = c
        self.h = C.c_void_p()
        rc = self.lib.dlt_lio_create(C.byref(c), C.byref(self.h))
        if rc != 0:
            self.h = None
            raise DltError(f"dlt_lio_create failed: {_ERR.get(rc, rc)}")
        self.out = LioScanOut()
        self.collect_after_scan = True
        # a non-owning view of the device handle (map export, neighbour read-back, ...)
        self.device = ScanToMap.__new__(ScanToMap)
        self.device.lib = self.lib
        self.device.h = C.c_void_p(self.lib.dlt_lio_device(self.h))
        self.device.close = lambda: None

    def close(self):
        if getattr(self, "h", None):
            if getattr(self, "_nccl", None):
                self.lib.dlt_sync(self.device.h)
                self.lib.dlt_nccl_destroy(self._nccl)
                self._nccl = None
            self.lib.dlt_lio_destroy(self.h)
            self.h = None
            self.device.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            msg = self.lib.dlt_lio_last_error(self.h)
            raise DltError(f"{_ERR.get(rc, rc)}: {msg.decode() if msg else ''}")

    def on_lidar_msg(self):
        self._ck(self.lib.dlt_lio_on_lidar_msg(self.h))

    def on_edge_count(self, n):
        self._ck(self.lib.dlt_lio_on_edge_count(self.h, C.c_int(n)))

    def force_imu_ready(self, mean_acc, last_imu7):
        a, b = _f64(mean_acc), _f64(last_imu7)
        self._ck(self.lib.dlt_lio_force_imu_ready(self.h, _p(a), _p(b)))

    def get_state(self) -> np.ndarray:
        s = np.zeros(612, np.float64)
        self._ck(self.lib.dlt_lio_get_state(self.h, _p(s)))
        return s

    def set_state(self, s, also_last=True):
        s = _f64(s)
        assert s.size == 612
        self._ck(self.lib.dlt_lio_set_state(self.h, _p(s), C.c_int(1 if also_last else 0)))

    def flags(self):
        f = np.zeros(8, np.int32)
        self._ck(self.lib.dlt_lio_get_flags(self.h, _p(f)))
        return dict(ekf_stop=int(f[0]), ekf_inited=int(f[1]), threshold=int(f[2]), lidar_cnt=int(f[3]), localmap_init=int(f[4]),
                    queue=int(f[5]), map_built=int(f[6]), imu_ready=int(f[7]))

    def localmap(self):
        b = np.zeros(6, np.float32)
        self._ck(self.lib.dlt_lio_get_localmap(self.h, _p(b)))
        return b

    def process_scan(self, pts48, lidar_beg_time, imu7, thermal: LioThermal | None = None) -> LioScanOut:
        """pts48 / imu7 may be numpy arrays or torch CPU tensors (pinned or not): only their data pointer is used."""
        if hasattr(pts48, "data_ptr"):
            n = int(pts48.shape[0])
            pp = C.c_void_p(pts48.data_ptr())
        else:
            a = np.ascontiguousarray(pts48, dtype=np.float32).reshape(-1, 12)
            n = a.shape[0]
            pp = _p(a)
        im = _f64(imu7).reshape(-1, 7)
        th = C.byref(thermal) if thermal is not None else None
        self._ck(self.lib.dlt_lio_process_scan(self.h, pp, C.c_int(n), C.c_double(lidar_beg_time), _p(im), C.c_int(im.shape[0]), th,
                                               C.byref(self.out)))
        return self._collect()

    def collect_insert(self):
        """async_insert: wait for the last scan's map_incremental, return its (downsample adds, raw adds)"""
        a, b = C.c_int(0), C.c_int(0)
        self._ck(self.lib.dlt_lio_collect_insert(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def _collect(self):
        # The library's default leaves map_incremental running when process_scan returns (added = -1).  This wrapper waits for
        # it by default so that callers read the counts of THIS scan (tests); collect_after_scan = False keeps the call as the
        # library delivers it (bench.py).
        if self.out.added < 0 and self.collect_after_scan:
            self.out.n_added_ds, self.out.n_added_raw = self.collect_insert()
            self.out.added = self.out.n_added_ds + self.out.n_added_raw
        return self.out

    def prefetch_scan(self, pts48):
        """start uploading a scan (pinned torch CPU tensor or contiguous float32 numpy array that stays alive and unchanged) ahead of its
        process_scan call: the copy overlaps whatever the device is doing"""
        if hasattr(pts48, "data_ptr"):
            n, pp = int(pts48.shape[0]), C.c_void_p(pts48.data_ptr())
        else:
            assert pts48.dtype == np.float32 and pts48.flags["C_CONTIGUOUS"]
            n, pp = pts48.reshape(-1, 12).shape[0], _p(pts48)
        self._ck(self.lib.dlt_lio_prefetch_scan(self.h, pp, C.c_int(n)))

    def process_scan_dev(self, pts48_dev_ptr: int, n: int, lidar_beg_time, observation_end_time, imu7, thermal: LioThermal | None = None) -> LioScanOut:
        """the scan is already resident in device memory (pointer to n 48-byte records)"""
        im = _f64(imu7).reshape(-1, 7)
        th = C.byref(thermal) if thermal is not None else None
        self._ck(self.lib.dlt_lio_process_scan_dev(self.h, C.c_void_p(pts48_dev_ptr), C.c_int(n), C.c_double(lidar_beg_time),
                                                   C.c_double(observation_end_time), _p(im), C.c_int(im.shape[0]), th, C.byref(self.out)))
        return self._collect()

    def process_cloud(self, cloud, layout, sensor, header_stamp, imu7, point_filter_num=5, min_range=0.5, max_range=1000.0, thermal=None):
        """sensor PointCloud2 payload (uint8) -> front end on the device -> the per-scan update (dlt_lio_process_cloud)"""
        buf = np.ascontiguousarray(cloud, dtype=np.uint8).reshape(-1)
        lay = (C.c_int * 7)(*[int(v) for v in layout])
        im = _f64(imu7).reshape(-1, 7)
        th = C.byref(thermal) if thermal is not None else None
        code = ScanToMap.SENSORS[sensor] if isinstance(sensor, str) else int(sensor)
        ns = C.c_int(0)
        self._ck(self.lib.dlt_lio_process_cloud(self.h, _p(buf), C.c_int(buf.size // int(layout[0])), lay, C.c_int(code), C.c_int(point_filter_num),
                                                C.c_float(min_range), C.c_float(max_range), C.c_double(header_stamp), _p(im), C.c_int(im.shape[0]), th,
                                                C.byref(self.out), C.byref(ns)))
        return self._collect(), ns.value

    def set_allreduce(self, device: str = "cuda"):
        """Sharded map: sum over torch.distributed ranks whatever the library hands to the callback (the partial normal
        equations after every evaluation of the measurement model, the per-point map_incremental decisions) in place, on
        the current stream (NCCL for device="cuda"; gloo on the CPU emulator build for device="cpu")."""
        import torch
        import torch.distributed as dist

        class _DevPtr:  # wrap a raw device pointer as a tensor without copying
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}

        cpu = device == "cpu"
        views = {}  # (pointer, n) -> tensor view: the library hands over the same few buffers every scan

        def _cb(ctx, buf, n):
            try:
                if dist.is_initialized() and dist.get_world_size() > 1:
                    t = views.get((buf, n))
                    if t is None:
                        if cpu:
                            t = torch.from_numpy(np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_double)), shape=(n,)))
                        else:
                            t = torch.as_tensor(_DevPtr(buf, n), device=device)
                        if len(views) < 64:
                            views[(buf, n)] = t
                    dist.all_reduce(t, op=dist.ReduceOp.SUM)
                return 0
            except Exception:  # never let an exception cross the C ABI
                import traceback

                traceback.print_exc()
                return 1

        self._red_cb = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int)(_cb)
        self._ck(self.lib.dlt_lio_set_reduce(self.h, self._red_cb, None, None))

    def attach_native_nccl(self, unique_id: bytes, rank: int, world: int, device: int):
        """Sharded map with the library's own NCCL transport (include/daliti_b200_nccl.h): ncclAllReduce(sum, double) enqueued on
        the handle's stream from C, no Python on the data path.  unique_id: native_nccl_unique_id() of rank 0, handed to
        every rank by the caller."""
        comm = C.c_void_p()
        idb = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        self.lib.dlt_nccl_last_error.restype = C.c_char_p
        if self.lib.dlt_nccl_create(idb, C.c_int(rank), C.c_int(world), C.c_int(device), C.byref(comm)) != 0:
            raise DltError("dlt_nccl_create: " + (self.lib.dlt_nccl_last_error() or b"").decode())
        if self.lib.dlt_nccl_attach(comm, self.h) != 0:
            raise DltError("dlt_nccl_attach failed")
        self._nccl = comm

    def native_nccl_unique_id(self) -> bytes:
        idb = (C.c_ubyte * 128)()
        self.lib.dlt_nccl_last_error.restype = C.c_char_p
        if self.lib.dlt_nccl_unique_id(idb) != 0:
            raise DltError("dlt_nccl_unique_id: " + (self.lib.dlt_nccl_last_error() or b"").decode())
        return bytes(idb)

    def peer_export(self) -> bytes:
        from .binding import PEER_BLOB_BYTES

        blob = (C.c_ubyte * PEER_BLOB_BYTES)()
        self._ck(self.lib.dlt_lio_peer_export(self.h, blob))
        return bytes(blob)

    def peer_attach(self, blobs):
        """Sharded map over NVLink peer memory: blobs = peer_export() of all dev.shard_count ranks in rank order.  Afterwards
        the sums over the ranks happen inside the kernels (no reduce callback, no collective launch)."""
        raw = b"".join(blobs)
        buf = (C.c_ubyte * len(raw)).from_buffer_copy(raw)
        self._ck(self.lib.dlt_lio_peer_attach(self.h, buf))

    def peer_detach(self):
        self._ck(self.lib.dlt_lio_peer_detach(self.h))

    def iters(self):
        n = self.out.n_iters
        arr = (LioIter * max(n, 1))()
        self.lib.dlt_lio_get_iters(self.h, arr, C.c_int(n))
        return [arr[i] for i in range(n)]

    def imu_poses(self) -> np.ndarray:
        out = np.zeros((512, 22), np.float64)
        n = self.lib.dlt_lio_get_imu_poses(self.h, _p(out), C.c_int(512))
        return out[:n]
