#!/usr/bin/env python
"""bench.py -- IEKF scan-to-map throughput / latency on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle/_ref + restated loop)

A step = the whole per-scan update of one synthetic LiDAR scan through the reference-facing
call (dlt_lio_process_scan): IMU deskew -> VoxelGrid -> 4 IEKF iterations {transform, (re)match
k=5, plane fit, residual, Jacobian, H^T H / H^T r reduction, 24-state solve} -> degeneracy
eigen-decomposition -> map_incremental.  Workload at N=1: BASELINE.json configs[1] (synthetic
OS1-64 scan, ~2 M-point map, IMU deskew, scan replay on one B200).  N>1: one independent
sequence per GPU, no collective on the data path (configs[4] style), weak scaling.

`value` (points/s) is measured with every scan already resident in HBM; `e2e` goes through
the same call with the scans in pinned HOST memory (H2D of the 48-byte records and D2H of the
normal equations inside the timed region).  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "IEKF scan-to-map points/sec (p50 per-scan update latency in ms_p50)"
UNIT = "points/s"


# ------------------------------------------------------------------------------------------ workload
def build_workload(rank: int, n_scans: int, workload: str):
    """scene + map + n_scans raw scans with IMU, deterministic per rank"""
    from daliti_b200 import synth

    lm_kwargs = {}

    if workload == "c1":  # VLP-16, ~30 k raw points, 200 k-point local map
        scene = synth.make_box_world(half=110.0, n_boxes=60, seed=10 + rank, keep_clear=6.0)
        traj = synth.Trajectory(speed=1.0, yaw_rate=0.1, z0=1.5)
        spec = synth.VLP16
        map_pts = synth.sample_map(scene, seed=10 + rank, max_points=200_000)
        name = "C1: synthetic VLP-16 scan (16x1800) vs ~200k-pt map"
    elif workload == "c3":  # degenerate tunnel, 64 x 1024, small local-map cube so that lasermap_fov_segment box-deletes fire
        scene = synth.make_tunnel(length=400.0)
        traj = synth.Trajectory(speed=5.0, yaw_rate=0.0, x0=-150.0, z0=1.5)
        spec = synth.ScanSpec(64, 1024, (-22.5, 22.5), max_range=60.0)
        map_pts = synth.sample_map(scene, seed=30 + rank)
        name = "C3: synthetic degenerate tunnel (3x3x400 m, ~19k-pt surface map at 0.5 m), 64x1024 scan at 5 m/s, insert every scan, cube 40 m / DET_RANGE 13 m so box deletes fire"
        lm_kwargs = dict(cube_len=40.0, det_range=13.0)
    elif workload == "c5":  # one of the independent 32-beam sequences: 32 x 1024 scan, ~0.5 M-point map
        scene = synth.make_box_world(half=160.0, n_boxes=260, seed=50 + rank, keep_clear=8.0, height=(8.0, 40.0), max_size=25.0)
        traj = synth.Trajectory(speed=2.0, yaw_rate=0.15, z0=1.8)
        spec = synth.ScanSpec(32, 1024, (-22.5, 22.5), max_range=100.0)
        map_pts = synth.sample_map(scene, seed=50 + rank, region=(-85.0, 85.0, -85.0, 85.0))
        name = "C5: independent synthetic 32-beam sequences (32x1024 scans, ~0.5M-pt map each)"
    else:  # c2 (default): OS1-64, ~125 k raw points, ~2 M-point map
        scene = synth.make_box_world(half=300.0, n_boxes=900, seed=20 + rank, keep_clear=8.0, height=(12.0, 70.0), max_size=30.0)
        traj = synth.Trajectory(speed=2.0, yaw_rate=0.2, z0=1.8)
        spec = synth.ScanSpec(64, 2048, (-22.5, 22.5), max_range=120.0)
        map_pts = synth.sample_map(scene, seed=20 + rank, region=(-135.0, 135.0, -135.0, 135.0))
        name = "C2: synthetic OS1-64 scan (64x2048) vs ~2M-pt map, IMU deskew, scan replay"
    seq = synth.Sequence(scene, traj, spec, seed=100 + rank)
    try:
        from concurrent.futures import ProcessPoolExecutor

        workers = max(1, min(16, len(os.sched_getaffinity(0)) - 1, n_scans))
        if workers > 1:
            with ProcessPoolExecutor(workers) as ex:
                scans = list(ex.map(_gen_scan, [(seq, k) for k in range(n_scans)]))
        else:
            scans = [seq.scan(k) for k in range(n_scans)]
    except Exception:
        scans = [seq.scan(k) for k in range(n_scans)]
    return dict(name=name, seq=seq, map_pts=map_pts, scans=scans, lm_kwargs=lm_kwargs)


def _gen_scan(a):
    seq, k = a
    return seq.scan(k)


def initial_state(seq):
    from daliti_b200 import synth

    s = np.zeros(612)
    s[0:24] = seq.traj.pose24(seq.t_start)
    s[24:27] = seq.traj.vel(seq.t_start)
    s[33:36] = [0, 0, -9.801]
    s[36:] = np.eye(24).ravel()
    last_imu = np.concatenate([[seq.t_start - 0.005], [0, 0, synth.G], [0, 0, seq.traj.yaw_rate]])
    return s, [0.0, 0.0, synth.G], last_imu


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML in a thread every 5 ms (the timed region of a
    50-scan run is only ~20-50 ms), falling back to an `nvidia-smi -lms` child when NVML is unavailable."""

    _BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index: int, period_s: float = 0.005):
        self.idx = gpu_index
        self.period = float(os.environ.get("BENCH_CLOCK_PERIOD", period_s))
        self.sm, self.mx, self.reasons = [], [], set()
        self.proc = None
        self.thread = None
        self.stop_flag = False
        self.mode = None

    def start(self):
        if self.period <= 0:
            return
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.idx
            if vis:
                try:
                    phys = int(vis.split(",")[self.idx])
                except Exception:
                    phys = self.idx
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
            self.mode = "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.mode = None
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.mode = "smi"
            self.thread = threading.Thread(target=self._read_smi, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in self._BAD.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(self.period)

    def _read_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            if len(r) >= 9 and r[1].replace(".", "").isdigit():
                self.sm.append(float(r[1]))
                if r[2].replace(".", "").isdigit():
                    self.mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)

    def stop(self):
        if self.mode is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        self.stop_flag = True
        if self.mode == "smi":
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        else:
            self.thread.join(timeout=1)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.mode}


# ------------------------------------------------------------------------------------------ CPU reference arm
class _QuietStdout:
    """the reference ikd-Tree printf()s from its constructor / rebuild thread: keep fd 1 clean for the JSON line"""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *a):
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved)
        os.close(self.null)


def run_cpu(work, n_steps, n_warm, threads_list):
    with _QuietStdout():
        return _run_cpu(work, n_steps, n_warm, threads_list)


def run_cpu_concurrent(works, n_steps, n_warm, threads_each):
    """N independent CPU pipelines at once (the reference arm's counterpart of N GPUs running one sequence each): one Python
    thread per pipeline (ctypes releases the GIL inside the oracle), threads_each OpenMP threads in each.  Aggregate = total
    points / wall time of the slowest pipeline."""
    import threading

    res = [None] * len(works)

    def one(i):
        res[i] = _run_cpu(works[i], n_steps, n_warm, [threads_each])

    with _QuietStdout():
        ths = [threading.Thread(target=one, args=(i,)) for i in range(len(works))]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        wall = time.perf_counter() - t0
    if any(r is None for r in res):
        raise RuntimeError("a CPU pipeline failed")
    tot_pts = sum(r["points_per_s"] * r["ms_per_step"] * 1e-3 * r["steps"] for r in res)
    slowest = max(r["ms_per_step"] * r["steps"] * 1e-3 for r in res)
    best = dict(res[0])
    best.update(points_per_s=tot_pts / slowest, ms_per_step=max(r["ms_per_step"] for r in res), ms_p50=float(np.median([r["ms_p50"] for r in res])),
                pipelines=len(works), wall_s=wall)
    return best


def _run_cpu(work, n_steps, n_warm, threads_list):
    """the reference's CPU path: unmodified ikd-Tree (oracle/_ref) under the restated loop (oracle/oracle.cpp)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob

    orc = ob.load()
    kind = ob.MAP_REF if orc.ref_ok else ob.MAP_PORT
    seq = work["seq"]
    best = None
    for thr in threads_list:
        lio = orc.new_lio(ob.default_lio_config(featptsThreshold=30), kind)
        s0, mean_acc, last_imu = initial_state(seq)
        lio.force_imu_ready(mean_acc, last_imu)
        lio.set_state(s0)
        lio.set_threads(thr)
        t_build = time.perf_counter()
        lio.map().build(work["map_pts"])
        t_build = time.perf_counter() - t_build
        times, pts_total, stages = [], 0, np.zeros(7)
        scans = work["scans"]
        for k in range(min(len(scans), n_warm + n_steps)):
            pts, t_beg, imu = scans[k]
            lio.on_lidar_msg()
            t0 = time.perf_counter()
            s = lio.process_scan(pts, t_beg, imu)
            dt = time.perf_counter() - t0
            if k >= n_warm:
                times.append(dt)
                pts_total += s.n_raw
                stages += np.array([s.t_deskew, s.t_voxel, s.t_knn, s.t_resid, s.t_solve, s.t_insert, s.t_delete])
        lio.close()
        tot = float(np.sum(times))
        res = dict(threads=thr, points_per_s=pts_total / tot if tot > 0 else 0.0, ms_per_step=1e3 * tot / max(1, len(times)),
                   ms_p50=1e3 * float(np.median(times)) if times else None, steps=len(times), map_build_s=t_build,
                   stage_ms=dict(zip(["deskew", "voxel", "knn", "resid_jacobian", "solve", "insert", "delete"],
                                     (1e3 * stages / max(1, len(times))).round(3).tolist())),
                   kind="reference" if kind == ob.MAP_REF else "port")
        if best is None or res["points_per_s"] > best["points_per_s"]:
            best = res
    return best


def run_parity(work, n_scans, threads, local_rank=0):
    """The device path against the oracle on the benchmark's own scans and map (tests/parity_tools.py; test infrastructure):
    identical inputs (the oracle's feats_down / poses / map through dlt_measure) and the whole pipeline re-synced per scan."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    from daliti_b200.binding import load_library
    from parity_tools import ParityRun

    with _QuietStdout():
        run = ParityRun(load_library(), ob.load(), work["seq"], work["map_pts"], lm_kwargs=work["lm_kwargs"], threads=threads, device=local_rank,
                        featptsThreshold=30)
        for k in range(n_scans):
            run.step(k, work["scans"][k])
        rec = run.summary()
        run.close()
    rec["ok"] = bool(rec["knn_sets_equal"] and rec["selected_equal"] and rec["effct_equal"] and rec["add_lists_equal"] and rec["map_contents_equal"]
                     and rec["n_iters_equal"] and rec["max_pose_rel_err"] < 1e-5 and rec["max_HtH_rel_err"] < 1e-10)
    rec["tolerances"] = {"pose_rel": 1e-5, "HtH_rel": 1e-10, "neighbour_sets / selections / counts / map contents": "exact"}
    rec["what"] = ("identical inputs: oracle feats_down + per-iteration pose + map-before through dlt_measure (knn_sets_equal, selected_equal, effct_equal, "
                   "max_HtH_rel_err, add_lists_equal, map_contents_equal); pipeline: dlt_lio_process_scan from the oracle's state and map, one scan at "
                   "a time (max_pose_rel_err, voxels_differing, pipeline_*)")
    return rec


# ------------------------------------------------------------------------------------------ main
_REAL_STDOUT = None


def _claim_stdout():
    """Everything libraries print (NCCL_DEBUG=VERSION banners, the reference tree's printf) goes to stderr; the one JSON
    line is written to the saved descriptor by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    txt = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(txt.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, txt)


REPLAY_SCANS = 100  # BASELINE.json configs[1] is a 100-scan replay: latency percentiles are taken over that many scans


def base_config(work):
    """the `config` object: the same keys and values in both arms (everything run-specific goes to `detail`)"""
    return {"workload": work["name"], "iterations": 4, "map_points": int(len(work["map_pts"])),
            "l2": "GPU arm: L2 flushed between timed steps (256 MiB memset outside the per-step CUDA-event pair)"}


def pin_rank(local_rank, world):
    """One block of host cores per rank: N replicas that spin on their own result flags must not share cores (r1: 0.92 weak
    scaling at N = 8 with zero communication was host contention).  Returns the cores or None."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = len(cores) // max(1, world)
        if world <= 1 or per < 1:
            return None
        mine = cores[local_rank * per:(local_rank + 1) * per]
        os.sched_setaffinity(0, mine)
        return mine
    except Exception:
        return None


def git_head():
    try:
        return subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True, timeout=5).stdout.strip() or None
    except Exception:
        return None


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--seqs-per-gpu", type=int, default=8, help="c5: independent sequences driven concurrently on every GPU (one stream + host thread each)")
    ap.add_argument("--c5-threads", type=int, default=0, help="c5: native worker threads of dlt_lio_replay_sequences per GPU (0 = min(sequences, host cores of the rank))")
    ap.add_argument("--device-loop", type=int, default=-1, choices=[-1, 0, 1, 2],
                    help="-1: the library's default (host loop); 1: iteration loop, zeta blend and map insert resident on the device (one sync "
                         "per scan); 2: loop on the device, blend/insert host-driven; 0: one host round trip per iteration")
    ap.add_argument("--shard-exchange", default="peer", choices=["peer", "nccl"],
                    help="c4: how the partial normal equations / map_incremental decisions are summed over the ranks: inside the kernels "
                         "over NVLink peer memory (dlt_peer_attach; falls back to nccl when the mailboxes cannot be mapped) or an NCCL all-reduce callback")
    ap.add_argument("--tiles", type=int, default=5, help="c4: the map is tiles x tiles shifted copies of the C2 map")
    ap.add_argument("--cpu-sample", type=int, default=3, help="scans of the same workload timed on the host cores (cpu_baseline) and checked for parity")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c4", action="store_true", help="N > 1: skip the sharded-map (C4) leg that rides along with the replica measurement")
    ap.add_argument("--c4-steps", type=int, default=40)
    ap.add_argument("--no-replay", action="store_true", help="N = 1: skip the 100-scan replay that the latency percentiles are taken from")
    args = ap.parse_args()
    K, W = args.steps, max(3, args.warmup)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        n_pipe = max(1, args.gpus)  # one CPU pipeline per GPU of our arm: the same weak-scaling job on the host cores
        works = [build_workload(i, K + W, args.workload) for i in range(n_pipe)]
        work = works[0]
        ncpu = len(os.sched_getaffinity(0))
        cand = sorted({1, 4, min(ncpu, 16)})
        # pick the thread count on two scans, then time K steps with it
        probe = {t: run_cpu(dict(work, scans=work["scans"][:3]), 2, 1, [t])["points_per_s"] for t in cand}
        thr = max(probe, key=probe.get)
        if n_pipe == 1:
            r = run_cpu(work, K, W, [thr])
        else:
            thr = max(1, min(thr, ncpu // n_pipe))
            r = run_cpu_concurrent(works, K, W, thr)
        line = {
            "impl": "reference", "metric": METRIC, "value": r["points_per_s"], "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"], "warmup": W,
            "ms_per_step": r["ms_per_step"], "ms_p50": r["ms_p50"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 geometry / f64 normal equations", "data": "synthetic",
            "config": base_config(work),
            "detail": {"threads_probe_points_per_s": {str(k): v for k, v in probe.items()}, "pipelines": n_pipe, "threads_per_pipeline": thr, "host_cores": ncpu,
                       "note": "N > 1: N independent CPU pipelines run concurrently on the box's host cores (one per GPU of the other arm)"},
            "cpu_baseline": {"value": r["points_per_s"], "unit": UNIT, "cores": thr * n_pipe, "kind": r["kind"],
                             "sample": f"{r['steps']} scans of the workload after {W} warm-up scans, {n_pipe} pipeline(s) x {thr} thread(s); as-shipped is 1 thread (both OpenMP pragmas commented out, laserMapping.cpp:827-828,946-947)",
                             "stage_ms": r["stage_ms"], "map_build_s": r["map_build_s"]},
            "e2e": {"value": r["points_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        emit(line)
        return 0

    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: daliti_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        if args.workload == "c5":
            # independent sequences: no data-path collective, and the ranks only meet for the start barrier and the final sums --
            # done over gloo, because an initialised NCCL communicator slows concurrent host threads of the same process down
            # by orders of magnitude here (18 ms instead of 0.36 ms per scan with 2 worker threads, measured; 1 thread is unaffected)
            dist.init_process_group("gloo")

        else:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from daliti_b200.lio import LaserMapping

    if args.workload == "c4":
        line = measure_c4(args, K, W, rank, local_rank, world, dist)
        if rank == 0:
            emit(line)
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0
    if args.workload == "c5":
        return main_c5(args, K, W, rank, local_rank, world, dist)

    replay = 0 if (args.no_replay or world > 1 or args.workload != "c2" or K >= REPLAY_SCANS) else REPLAY_SCANS
    work = build_workload(rank, max(K, replay) + W, args.workload)
    k4 = max(10, min(args.c4_steps, K))
    work_c4 = None
    if world > 1 and not args.no_c4:  # the sharded-map leg: every rank cooperates on the SAME scans (rank 0's)
        work_c4 = work if (rank == 0 and args.workload == "c2" and len(work["scans"]) >= k4 + W) else build_workload(0, k4 + W, "c2")
    # ---- N > 1: the sharded-map configuration (C4) rides along, so that the scaling record holds a partitioned curve too.  It runs
    # FIRST, in the state a sharded job has: every rendezvous of the sharded ranks waits for the slowest host thread, and behind
    # the replica legs (per-rank core blocks, their threads and allocations) the same configuration measured 0.38-0.39 ms at 8 GPUs
    # against 0.243 ms in a process of its own (profiles/r02_g8_bench_default.json, r02_g8_default_b.json, r02_g8_c4_host.json).
    c4_line, c4_err = None, None
    if world > 1 and not args.no_c4:
        try:
            c4_line = measure_c4(args, k4, W, rank, local_rank, world, dist, work=work_c4)
        except Exception as e:
            c4_err = repr(e)
    pinned_cores = None if os.environ.get("BENCH_NO_PIN") else pin_rank(local_rank, world)
    seq, scans = work["seq"], work["scans"]
    n_map = len(work["map_pts"])
    stream = torch.cuda.Stream(device=local_rank)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")

    def reset(device_loop=args.device_loop):
        s0, mean_acc, last_imu = initial_state(seq)
        lm2 = LaserMapping(dev=dict(device=local_rank, max_scan_points=1 << 18, max_map_points=max(1 << 22, 2 * n_map)), featptsThreshold=30,
                           device_loop=device_loop, **work["lm_kwargs"])
        lm2.collect_after_scan = False  # the call as the library delivers it: map_incremental keeps running on its own stream (async_insert)
        lm2.device.set_stream(stream.cuda_stream)
        lm2.force_imu_ready(mean_acc, last_imu)
        lm2.set_state(s0)
        lm2.device.map_build(work["map_pts"])
        return lm2

    # device-resident copies (for `value`) and pinned host copies (for `e2e`) of every scan
    dev_scans, pin_scans = [], []
    for pts, t_beg, imu in scans:
        t = torch.from_numpy(np.ascontiguousarray(pts))
        pin_scans.append(t.pin_memory())
        dev_scans.append(t.to(f"cuda:{local_rank}"))
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def run(mode, flush, K=K):
        """W warm-up scans, then K timed scans.  Returns per-step ms (CUDA events on the launching stream), outputs."""
        lmx = reset()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        outs, host_ms = [], []
        launches0 = None
        # no collector pauses between a step's start event and its first launch: a generation-2 collection of this script's own
        # objects landed inside the same step of every run (1-4 ms of idle GPU inside one event pair; the host clock around the
        # call itself showed nothing)
        gc.collect()
        gc.disable()
        with torch.cuda.stream(stream):
            for k in range(W + K):
                pts, t_beg, imu = scans[k]
                lmx.on_lidar_msg()
                if k == W:
                    barrier()
                    launches0 = lmx.device.launch_count()
                if flush:
                    flush_buf.zero_()  # evict L2 (126 MB) between steps; outside the per-step event pair
                if k >= W:
                    ev[k - W][0].record(stream)
                    th0 = time.perf_counter()
                if mode == "dev":
                    o = lmx.process_scan_dev(dev_scans[k].data_ptr(), len(pts), t_beg, t_beg + float(pts[-1, 6]), imu)
                else:
                    if mode == "host_pf" and k + 1 < W + K:
                        lmx.prefetch_scan(pin_scans[k + 1])  # double buffering: scan k+1 crosses PCIe while scan k is processed
                    o = lmx.process_scan(pin_scans[k], t_beg, imu)
                if k >= W:
                    ev[k - W][1].record(stream)
                    host_ms.append(1e3 * (time.perf_counter() - th0))
                    outs.append((o.n_raw, o.n_down, o.n_iters, o.ekf_stop, o.added, o.map_points_before, lmx.iters()[-1].effct_feat_num if o.n_iters else 0,
                                 o.t_deskew, o.t_voxel, o.t_iterate, o.t_insert, o.t_delete, o.t_total, o.deleted, o.degenerate))
            barrier()
            launches = lmx.device.launch_count() - launches0
        gc.enable()
        ms = np.array([a.elapsed_time(b) for a, b in ev])
        return lmx, ms, np.array(host_ms), outs, launches

    sampler = ClockSampler(local_rank)
    sampler.start()
    lm_v, ms_v, host_v, outs_v, launches = run("dev", flush=True)
    clocks = sampler.stop()
    lm_v.close()
    lm_w, ms_w, _, _, _ = run("dev", flush=False)
    lm_w.close()
    lm_s, ms_s, _, outs_s, _ = run("host", flush=True)       # serial: upload, then update, inside every step
    lm_s.close()
    lm_e, ms_e, host_e, outs_e, _ = run("host_pf", flush=True)  # the upload of scan k+1 overlaps the update of scan k
    lm_e.close()
    replay_rec = None
    if replay:  # latency percentiles over a 100-scan replay (the K timed steps above give `value`)
        lm_r, ms_r, _, outs_r, _ = run("dev", flush=True, K=replay)
        lm_r.close()
        lm_r, ms_rs, _, _, _ = run("host", flush=True, K=replay)
        lm_r.close()
        replay_rec = {"scans": replay, "ms_p50": float(np.median(ms_r)), "ms_p90": float(np.percentile(ms_r, 90)), "ms_p99": float(np.percentile(ms_r, 99)),
                      "ms_max": float(ms_r.max()), "ms_mean": float(ms_r.mean()), "serial_e2e_ms_p50": float(np.median(ms_rs)),
                      "serial_e2e_ms_p99": float(np.percentile(ms_rs, 99)), "iterations_run_mean": float(np.mean([o[2] for o in outs_r])),
                      "note": "device-resident scans, L2 flushed between scans, per-scan CUDA events; serial_e2e: dlt_lio_process_scan from pinned host buffers (upload, then update)"}

    # ---- per-kernel device time (separate short pass with event pairs around each kernel group)
    prof = None
    roof = None
    try:
        n_prof, nd_sum, nraw_sum, match_passes, iters_sum = 0, 0, 0, 0, 0
        lm_p = reset(device_loop=0)  # one real launch per event pair (the device-resident loop also enqueues no-op launches)
        lm_p.device.set_profiling(True)
        with torch.cuda.stream(stream):
            for k in range(min(W + K, W + 12)):
                pts, t_beg, imu = scans[k]
                lm_p.on_lidar_msg()
                if k == W:
                    lm_p.device.get_profile(reset=True)
                flush_buf.zero_()
                o = lm_p.process_scan_dev(dev_scans[k].data_ptr(), len(pts), t_beg, t_beg + float(pts[-1, 6]), imu)
                if k >= W:
                    n_prof += 1
                    nd_sum += o.n_down
                    nraw_sum += o.n_raw
                    iters_sum += o.n_iters
                    match_passes += sum(1 for it in lm_p.iters() if it.did_match)
        prof = lm_p.device.get_profile(reset=True)
        lm_p.close()
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peaks = json.load(open(pk))
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        knn_ms, knn_n = prof["knn"]      # whole match pass: k_knn8 + k_knn (+ two 4-byte memsets)
        k8_ms, k8_n = prof["knn8"]       # k_knn8 alone
        res_ms, res_n = prof["residual"]
        nd_mean = nd_sum / max(1, n_prof)
        nraw_mean = nraw_sum / max(1, n_prof)
        # algorithmic bytes per launch: kNN pass  N*(16 + 5*16 + 5*4), residual pass N*(16+16) + 92*8
        knn_bytes = nd_mean * (16 + 5 * 16 + 5 * 4)
        res_bytes = nd_mean * 32 + 92 * 8
        dom = "k_knn8" if k8_ms >= res_ms else "k_residual"
        d_ms, d_n, d_bytes = (k8_ms, k8_n, knn_bytes) if dom == "k_knn8" else (res_ms, res_n, res_bytes)
        achieved = d_bytes / (1e-3 * d_ms / max(1, d_n)) / 1e9 if d_ms > 0 else 0.0
        traffic, traffic_src, traffic_commit, inst_per_launch = None, None, None, None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")  # dram__bytes_read+write per launch from the committed --set full capture
        if os.path.exists(tp):
            tj = json.load(open(tp))
            if dom in tj:
                traffic, traffic_src = tj[dom].get("dram_bytes_per_launch"), tj[dom].get("source")
                traffic_commit, inst_per_launch = tj[dom].get("commit"), tj[dom].get("warp_instructions_per_launch")
        launch_us = 1e3 * d_ms / max(1, d_n)
        # whole-step figure (SURVEY.md 8d formula for B_scan) against the same peak
        mp_scan, it_scan = match_passes / max(1, n_prof), iters_sum / max(1, n_prof)
        b_scan = nraw_mean * 48 + nraw_mean * 16 + nd_mean * 16 + mp_scan * nd_mean * 116 + it_scan * nd_mean * 32 + it_scan * 92 * 8
        step_ms = float(ms_v.mean())
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        issue_peak = 148 * 4 * sm_mhz * 1e6  # warp instructions / s: 4 schedulers per SM, one instruction per cycle each
        issue = None
        if inst_per_launch:
            ia = inst_per_launch / (launch_us * 1e-6)
            issue = {"bound": "issue", "kernel": dom, "achieved": ia / 1e9, "peak": issue_peak / 1e9, "unit": "G warp-instructions/s", "frac": ia / issue_peak,
                     "warp_instructions_per_launch": inst_per_launch, "from_committed_profile": True, "profile_commit": traffic_commit,
                     "note": "instruction count from the committed ncu --set full capture, launch time measured in this run"}
        roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src, "traffic_from_committed_profile": traffic is not None, "traffic_profile_commit": traffic_commit,
                "peak_source": peak_src, "avg_launch_us": launch_us, "algorithmic_bytes_per_launch": d_bytes,
                "note": "single-scan working set is L2-resident and the kernel is issue/latency-bound (SURVEY.md 8d); the fraction is reported, not a target at this size",
                "issue": issue,
                "step": {"bound": "hbm", "achieved": b_scan / (step_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": b_scan / (step_ms * 1e-3) / 1e9 / peak,
                         "algorithmic_bytes_per_scan": b_scan, "ms_per_scan": step_ms, "match_passes_per_scan": mp_scan, "iterations_per_scan": it_scan},
                "kernel_ms_per_scan": {k: round(v[0] / max(1, n_prof), 4) for k, v in prof.items() if v[1]},
                "kernel_launch_groups_per_scan": {k: round(v[1] / max(1, n_prof), 2) for k, v in prof.items() if v[1]}}
    except Exception as e:  # profiling is auxiliary: never lose the headline number over it
        roof = {"bound": "hbm", "achieved": None, "peak": None, "unit": "GB/s", "frac": None, "traffic": None, "error": repr(e)}

    # ---- aggregate over ranks: total points / max time
    pts_v = float(sum(o[0] for o in outs_v))
    pts_e = float(sum(o[0] for o in outs_e))
    t_v, t_w, t_e, t_s = float(ms_v.sum()), float(ms_w.sum()), float(ms_e.sum()), float(ms_s.sum())
    if dist is not None:
        buf = torch.tensor([pts_v, pts_e, t_v, t_w, t_e, float(launches), t_s], dtype=torch.float64, device=f"cuda:{local_rank}")
        tot = buf.clone()
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        mx = buf.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        pts_v, pts_e, launches = float(tot[0]), float(tot[1]), int(tot[5])
        t_v, t_w, t_e, t_s = float(mx[2]), float(mx[3]), float(mx[4]), float(mx[6])

    cpu, parity = None, None
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        try:
            ncpu = len(os.sched_getaffinity(0))
            cwork = dict(work, scans=scans[: 1 + args.cpu_sample])
            cpu_r = run_cpu(cwork, args.cpu_sample, 1, sorted({1, min(4, ncpu)}))
            cpu = {"value": cpu_r["points_per_s"], "unit": UNIT, "cores": cpu_r["threads"], "kind": cpu_r["kind"],
                   "sample": f"{cpu_r['steps']} scans of the same workload after 1 warm-up scan (reference ikd-Tree + restated loop); best of 1 thread (as shipped) and 4 threads (the commented-out omp pragmas)",
                   "ms_per_scan": cpu_r["ms_per_step"], "stage_ms": cpu_r["stage_ms"], "map_build_s": cpu_r["map_build_s"]}
        except Exception as e:
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(e)}
        try:
            parity = run_parity(work, max(1, args.cpu_sample), min(8, len(os.sched_getaffinity(0))), local_rank)
        except Exception as e:
            parity = {"ok": False, "error": repr(e)}

    c4 = None
    if world > 1 and not args.no_c4:
        if c4_err is not None:
            c4 = {"error": c4_err}
        elif rank == 0 and c4_line is not None:
            c4 = {"ms_p50": c4_line["ms_p50"], "ms_per_step": c4_line["ms_per_step"], "ms_p99": c4_line["ms_p99"], "points_per_s": c4_line["value"],
                  "scans_per_s": c4_line["scans_per_s"], "steps": c4_line["steps"], "nranks": world, "scaling": "strong",
                  "exchange": c4_line["config"]["shard_exchange"], "loop": c4_line["config"]["loop"], "map_points": c4_line["config"]["map_points"],
                  "live_points_incl_halos": c4_line["config"]["live_points_incl_halos"], "map_build_s": c4_line["config"]["map_build_s"],
                  "e2e": c4_line["e2e"], "gpu_launches": c4_line["gpu_launches"], "workload": c4_line["config"]["workload"],
                  "unsharded_replica_ms_p50_same_run": float(np.median(ms_v)),
                  "order": "measured before the replica legs, on the process's full core set",
                  "sharded_over_unsharded_p50": c4_line["ms_p50"] / float(np.median(ms_v))}

    if rank == 0:
        n_raw_mean = float(np.mean([o[0] for o in outs_v]))
        n_down_mean = float(np.mean([o[1] for o in outs_v]))
        iters_mean = float(np.mean([o[2] for o in outs_v]))
        # per scan, device loop: the 48-byte records + IMU poses + the IEKF block prefix (state_propagat, thermal delta, state,
        # last_nodegared, window) up; the block's in/out + out part with 4 iteration records and the 8 map counters down
        if args.device_loop in (-1, 0):  # host loop: the normal equations come back once per iteration
            h2d = n_raw_mean * 48 + 22 * 8 * 22
            d2h = iters_mean * 160 * 8 + 48 + 42 * 8 + 2 * 32
        else:
            h2d = n_raw_mean * 48 + 22 * 8 * 22 + (36 + 36 + 1 + 2 * 612) * 8 + 28 * 4
            d2h = (2 * 612 + 42) * 8 + 24 * 4 + 4 * 1952 + 32
        line = {
            "metric": METRIC, "value": pts_v / (t_v * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": t_v / K, "ms_p50": float(np.median(ms_v)), "ms_p99": float(np.percentile(ms_v, 99)),
            "scans_per_s": world * K / (t_v * 1e-3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 geometry / f64 normal equations", "data": "synthetic",
            "config": base_config(work),
            "detail": {"iterations_run_mean": iters_mean,
                       "loop": {-1: "library default: host loop over dlt_measure, result block written straight into pinned host memory (no copy, no stream sync per iteration)",
                                1: "device-resident loop + zeta blend + map insert (dlt_iekf_update, one fused launch per iteration), 1 host sync per scan",
                                2: "device-resident loop (dlt_iekf_update), blend/insert host-driven, 2 host syncs per scan",
                                0: "host loop over dlt_measure"}[args.device_loop],
                       "deleted_total": int(sum(o[13] for o in outs_v)), "degenerate_scans": int(sum(o[14] for o in outs_v)),
                       "n_raw_mean": n_raw_mean, "n_down_mean": n_down_mean, "effct_feat_mean": float(np.mean([o[6] for o in outs_v])),
                       "ekf_stops": int(sum(o[3] for o in outs_v)), "l2": "flushed between steps (256 MiB memset outside the per-step CUDA-event pair)",
                       "timing": "sum of per-step CUDA-event times on the launching stream, max over ranks; Python's cyclic collector is off inside the timed loop (gc.collect() before it)",
                       "insert": "async_insert (library default): map_incremental of scan k runs on the handle's insert stream; the step's end event marks the final pose, the insert overlaps the next scan's deskew / VoxelGrid and its first match pass waits for it on the device",
                       "parallelism": "1 sequence per GPU, no data-path collective" if world > 1 else "single GPU",
                       "host_cores_of_this_rank": pinned_cores,
                       "value_l2_warm_points_per_s": pts_v / (t_w * 1e-3), "ms_p50_l2_warm": float(np.median(ms_w)),
                       "host_ms_p50": float(np.median(host_v)),
                       "ms_p90": float(np.percentile(ms_v, 90)), "ms_max": float(ms_v.max()),
                       "slow_steps": [{"step": int(i), "ms": round(float(ms_v[i]), 4), "iters": int(outs_v[i][2]), "n_down": int(outs_v[i][1]),
                                       "host_stage_ms": [round(1e3 * float(x), 4) for x in outs_v[i][7:13]]}
                                      for i in np.argsort(-ms_v)[:5] if ms_v[i] > 1.5 * np.median(ms_v)],
                       "host_stage_ms_mean": dict(zip(["deskew_enqueue", "voxelgrid", "iterations", "insert_and_eigen", "delete", "total"],
                                                      (1e3 * np.mean([o[7:13] for o in outs_v], axis=0)).round(4).tolist())),
                       "git": git_head()},
            "e2e": {"value": pts_e / (t_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": t_e / K, "ms_p50": float(np.median(ms_e)),
                    "api": "dlt_lio_prefetch_scan(scan k+1) + dlt_lio_process_scan(scan k), pinned host buffers: every scan's 48-byte records cross "
                           "PCIe inside the timed loop, overlapped with the previous scan's update (double-buffered upload)",
                    "serial": {"value": pts_e / (t_s * 1e-3), "ms_per_step": t_s / K, "ms_p50": float(np.median(ms_s)),
                               "note": "dlt_lio_process_scan alone: upload, then update, inside every step (message-to-pose latency)"}},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "parity": parity,
            "replay": replay_rec,
            "c4": c4,
        }
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and parity is not None and not parity.get("ok", False):
        print("bench.py: PARITY FAILED against the oracle: " + json.dumps(parity), file=sys.stderr)
        return 3
    return 0


def measure_c4(args, K, W, rank, local_rank, world, dist, work=None):
    """BASELINE config C4: a ~50 M-point map (tiles x tiles shifted copies of the C2 map) spatially sharded over the
    ranks; every rank evaluates the query points it owns and the 158-double normal equations are summed with one
    NCCL all-reduce per IEKF iteration (dlt_lio_set_reduce).  Poses hop from tile to tile, so consecutive scans
    touch different parts of the map.  Strong scaling: the work per scan is fixed.  Returns the JSON line on rank 0 (None elsewhere)."""
    import torch

    from daliti_b200.lio import LaserMapping

    if work is None:
        work = build_workload(0, K + W, "c2")  # every rank sees the same scans: they cooperate on each one
    seq, scans = work["seq"], work["scans"]
    base = work["map_pts"]
    pitch = 270.0
    T = args.tiles
    offs = [np.array([ix * pitch, iy * pitch, 0.0]) for ix in range(T) for iy in range(T)]
    n_map = len(base) * len(offs)
    per_rank_cap = int(n_map * (1.0 if world == 1 else min(1.0, 1.6 / world + 0.15))) + (1 << 20)
    lm = LaserMapping(dev=dict(device=local_rank, max_scan_points=1 << 18, max_map_points=per_rank_cap, shard_rank=rank, shard_count=world,
                               shard_tile_shift=5), featptsThreshold=30, device_loop=args.device_loop,
                      cube_len=1.0e6)  # the local-map cube must cover the whole tiled area: no lasermap_fov_segment deletes
    stream = torch.cuda.Stream(device=local_rank)
    lm.collect_after_scan = False  # the call as the library delivers it (async_insert)
    lm.device.set_stream(stream.cuda_stream)
    s0, mean_acc, last_imu = initial_state(seq)
    lm.force_imu_ready(mean_acc, last_imu)
    lm.set_state(s0)
    t_build = time.perf_counter()
    first = True
    for o in offs:  # dlt_map_build resets the map: build the first tile, add the others raw (Add_Points(.., false))
        tile = base.copy()
        tile[:, :3] += o.astype(np.float32)
        if first:
            lm.device.map_build(tile)
            first = False
        else:
            lm.device.map_add(tile, False)
    t_build = time.perf_counter() - t_build
    live = lm.device.map_valid_count()
    exchange = "none"
    if world > 1:
        if args.shard_exchange == "peer":
            from daliti_b200.sharded import attach_peers

            exchange = "peer" if attach_peers(lm) else "nccl (peer mailboxes could not be mapped)"
        if exchange != "peer":
            exchange = exchange if exchange != "none" else "nccl"
            try:  # the library's own transport: ncclAllReduce issued from C on the handle's stream (include/daliti_b200_nccl.h)
                from daliti_b200.sharded import attach_native_nccl

                attach_native_nccl(lm, local_rank)
                exchange += " (native: dlt_nccl_allreduce)"
            except Exception:
                with torch.cuda.stream(stream):
                    lm.set_allreduce(f"cuda:{local_rank}")
                exchange += " (torch.distributed callback)"
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")
    dev_scans = [torch.from_numpy(np.ascontiguousarray(p)).to(f"cuda:{local_rank}") for p, _, _ in scans]
    pin_scans = [torch.from_numpy(np.ascontiguousarray(p)).pin_memory() for p, _, _ in scans]

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def run(mode):
        lm.set_state(s0)
        cur = np.zeros(3)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        outs = []
        launches0 = 0
        gc.collect()
        gc.disable()
        with torch.cuda.stream(stream):
            for k in range(W + K):
                pts, t_beg, imu = scans[k]
                nxt = offs[(7 * k) % len(offs)]
                st = lm.get_state()
                st[9:12] += nxt - cur  # hop to another tile: the tiles are exact shifted copies
                cur = nxt
                lm.set_state(st, also_last=True)
                lm.on_lidar_msg()
                if k == W:
                    barrier()
                    launches0 = lm.device.launch_count()
                flush_buf.zero_()
                if k >= W:
                    ev[k - W][0].record(stream)
                if mode == "dev":
                    o = lm.process_scan_dev(dev_scans[k].data_ptr(), len(pts), t_beg, t_beg + float(pts[-1, 6]), imu)
                else:
                    if k + 1 < W + K:
                        lm.prefetch_scan(pin_scans[k + 1])  # double-buffered upload
                    o = lm.process_scan(pin_scans[k], t_beg, imu)
                if k >= W:
                    ev[k - W][1].record(stream)
                    outs.append((o.n_raw, o.n_down, o.n_iters, lm.iters()[-1].effct_feat_num if o.n_iters else 0))
            barrier()
        gc.enable()
        return np.array([a.elapsed_time(b) for a, b in ev]), outs, lm.device.launch_count() - launches0

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_v, outs_v, launches = run("dev")
    clocks = sampler.stop()
    ms_e, outs_e, _ = run("host")
    t_v, t_e = float(ms_v.sum()), float(ms_e.sum())
    if dist is not None:
        buf = torch.tensor([t_v, t_e, float(launches), float(live)], dtype=torch.float64, device=f"cuda:{local_rank}")
        mx = buf.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = buf.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        t_v, t_e, launches, live_sum = float(mx[0]), float(mx[1]), int(sm[2]), int(sm[3])
    else:
        live_sum = live
    if os.environ.get("BENCH_C4_TIMELINE"):  # development aid: device timeline of two more scans (CUDA events around every kernel group)
        lm.device.set_profiling(True)
        with torch.cuda.stream(stream):
            for k in range(3):
                pts, t_beg, imu = scans[W + k]
                lm.on_lidar_msg()
                lm.device.get_profile(reset=True)
                barrier()
                t0 = time.perf_counter()
                o = lm.process_scan_dev(dev_scans[W + k].data_ptr(), len(pts), t_beg, t_beg + float(pts[-1, 6]), imu)
                host_ms = 1e3 * (time.perf_counter() - t0)
                tl = lm.device.get_timeline()
                if k == 2:
                    txt = [f"rank {rank}: host {host_ms:.3f} ms  stages deskew {1e3*o.t_deskew:.3f} voxel {1e3*o.t_voxel:.3f} iterate {1e3*o.t_iterate:.3f} insert {1e3*o.t_insert:.3f}"]
                    prev = 0.0
                    for name, a, b in tl:
                        if name == "knn8":
                            continue
                        txt.append(f"  rank {rank} {name:10s} start {1e3*a:7.1f} dur {1e3*(b-a):6.1f} gap {1e3*(a-prev):6.1f}")
                        if name != "eigen6":
                            prev = b
                    print("\n".join(txt), file=sys.stderr, flush=True)
        lm.device.set_profiling(False)
    line = None
    if rank == 0:
        pts_total = float(sum(o[0] for o in outs_v))
        loop_txt = {-1: "library default (host loop over dlt_measure; with peers the sum over the ranks happens inside k_residual)",
                    0: "host loop over dlt_measure", 1: "device-resident loop + blend + insert", 2: "device-resident loop, blend / insert host-driven"}[args.device_loop]
        line = {
            "metric": METRIC, "value": pts_total / (t_v * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": t_v / K, "ms_p50": float(np.median(ms_v)), "ms_p99": float(np.percentile(ms_v, 99)), "scans_per_s": K / (t_v * 1e-3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 geometry / f64 normal equations", "data": "synthetic",
            "config": {"workload": f"C4: {n_map / 1e6:.1f}M-pt voxel-hash map ({T}x{T} tiles of the C2 map) spatially sharded over {world} GPU(s), "
                                   "C2 scans at poses hopping between tiles, H^T H / H^T r summed over the ranks per iteration, map_incremental with the owners' decisions exchanged",
                       "iterations": 4, "n_raw_mean": float(np.mean([o[0] for o in outs_v])), "n_down_mean": float(np.mean([o[1] for o in outs_v])),
                       "map_points": n_map, "live_points_incl_halos": live_sum, "map_build_s": t_build, "loop": loop_txt,
                       "effct_feat_mean": float(np.mean([o[3] for o in outs_v])),
                       "l2": "flushed between steps (256 MiB memset outside the per-step CUDA-event pair)",
                       "timing": "sum of per-step CUDA-event times on the launching stream, max over ranks; Python's cyclic collector is off inside the timed loop (gc.collect() before it)",
                       "insert": "async_insert (library default): map_incremental of scan k runs on the handle's insert stream; the step's end event marks the final pose, the insert overlaps the next scan's deskew / VoxelGrid and its first match pass waits for it on the device",
                       "parallelism": (f"map sharded {world}-way by 32-cell tiles + halo; 158 doubles summed over the ranks per iteration "
                                       + ("inside k_residual through NVLink peer mailboxes (CUDA IPC): no collective launch"
                                          if exchange == "peer" else "by an NCCL all-reduce between k_residual and k_iekf_step")) if world > 1 else "single GPU, unsharded",
                       "shard_exchange": exchange},
            "e2e": {"value": float(sum(o[0] for o in outs_e)) / (t_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(np.mean([o[0] for o in outs_e]) * 48 + 22 * 8 * 22),
                    "d2h_bytes_per_step": int(np.mean([o[2] for o in outs_e]) * (158 * 8 + 4) + 48 + 42 * 8), "ms_per_step": t_e / K,
                    "ms_p50": float(np.median(ms_e)), "ms_p99": float(np.percentile(ms_e, 99)),
                    "api": "dlt_lio_prefetch_scan(next) + dlt_lio_process_scan (pinned host buffers) + " + ("dlt_lio_peer_attach" if exchange == "peer" else "dlt_lio_set_reduce")},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": None, "cpu_baseline": None,
        }
    lm.close()
    del dev_scans, pin_scans, flush_buf
    torch.cuda.empty_cache()
    return line


def main_c5(args, K, W, rank, local_rank, world, dist):
    """BASELINE config C5: independent 32-beam sequences, S per GPU replayed concurrently by the library's native multi-sequence
    driver (dlt_lio_replay_sequences: one handle + CUDA streams per sequence, worker threads that multiplex sequences as
    coroutines), no data-path collective.  A step = one scan of every sequence; value = aggregate raw points / s."""
    import torch

    from daliti_b200.lio import LaserMapping

    S = max(1, args.seqs_per_gpu)
    works = [build_workload(rank * S + i, K + W, "c5") for i in range(S)]
    dev = f"cuda:{local_rank}"
    streams, dev_scans, pin_scans = [], [], []
    for w in works:
        streams.append(torch.cuda.Stream(device=local_rank))
        dev_scans.append([torch.from_numpy(np.ascontiguousarray(p)).to(dev) for p, _, _ in w["scans"]])
        pin_scans.append([torch.from_numpy(np.ascontiguousarray(p)).pin_memory() for p, _, _ in w["scans"]])

    from daliti_b200.lio import replay_sequences

    n_thr = args.c5_threads if args.c5_threads > 0 else max(1, min(S, len(os.sched_getaffinity(0))))

    def run(mode, K=K):
        """every sequence replays W warm-up scans, then K timed scans, through dlt_lio_replay_sequences: native worker threads,
        several sequences per thread as coroutines that switch wherever the library waits for the device"""
        lms = []
        for w, st in zip(works, streams):
            lm = LaserMapping(dev=dict(device=local_rank, max_scan_points=1 << 16, max_map_points=max(1 << 20, 2 * len(w["map_pts"])),
                                       voxel_bitmap_bits=1 << 25), featptsThreshold=30, device_loop=args.device_loop)
            lm.collect_after_scan = False
            lm.device.set_stream(st.cuda_stream)
            s0, mean_acc, last_imu = initial_state(w["seq"])
            lm.force_imu_ready(mean_acc, last_imu)
            lm.set_state(s0)
            lm.device.map_build(w["map_pts"])
            lms.append(lm)

        def lists(lo, hi):
            out = []
            for i, w in enumerate(works):
                if mode == "dev":
                    out.append([(dev_scans[i][k], w["scans"][k][1], w["scans"][k][2], w["scans"][k][1] + float(w["scans"][k][0][-1, 6])) for k in range(lo, hi)])
                else:
                    out.append([(pin_scans[i][k], w["scans"][k][1], w["scans"][k][2]) for k in range(lo, hi)])
            return out

        replay_sequences(lms, lists(0, W), n_threads=n_thr, prefetch=(mode != "dev"), want_outs=False)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        start = torch.cuda.Event(enable_timing=True)
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(S)]
        l0 = lms[0].device.launch_count()
        start.record(torch.cuda.current_stream())
        outs = replay_sequences(lms, lists(W, W + K), n_threads=n_thr, prefetch=(mode != "dev"))
        for e, st in zip(ends, streams):
            e.record(st)
        torch.cuda.synchronize()
        ms = max(start.elapsed_time(e) for e in ends)
        n_launch = lms[0].device.launch_count() - l0
        stats = []
        for lm, o in zip(lms, outs):
            lm.out.n_iters = o[-1].n_iters if o else 0  # (the wrapper's own record was not filled by the native driver)
            its = lm.iters()
            eff = its[-1].effct_feat_num if its else 0
            stats.append((sum(x.n_raw for x in o), sum(x.n_down for x in o), eff * len(o)))
        for lm in lms:
            lm.close()
        return ms, stats, n_launch

    sampler = ClockSampler(local_rank)
    sampler.start()
    # one untimed priming pass: with several ranks on a box the first multi-threaded pass of a process runs an order of
    # magnitude slower than every later one (whichever mode goes first, measured), far beyond what W warm-up scans absorb
    run("dev", K=min(K, 3))
    ms_v, st_v, launches = run("dev")
    clocks = sampler.stop()
    ms_e, st_e, _ = run("host")
    pts_v, pts_e = float(sum(s[0] for s in st_v)), float(sum(s[0] for s in st_e))
    if dist is not None:
        buf = torch.tensor([pts_v, pts_e, float(launches)], dtype=torch.float64)
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        mx = torch.tensor([ms_v, ms_e], dtype=torch.float64)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        pts_v, pts_e, launches = float(buf[0]), float(buf[1]), int(buf[2])
        ms_v, ms_e = float(mx[0]), float(mx[1])
    if rank == 0:
        n_seq = S * world
        line = {
            "metric": METRIC, "value": pts_v / (ms_v * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_v / K,
            "scans_per_s": n_seq * K / (ms_v * 1e-3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 geometry / f64 normal equations", "data": "synthetic",
            "config": {"workload": f"{works[0]['name']}: {n_seq} sequences, {S} per GPU replayed concurrently by dlt_lio_replay_sequences (one handle + CUDA streams each; {n_thr} native worker threads per GPU, sequences multiplexed as coroutines), "
                                   f"{K} scans per sequence timed (BASELINE names 1000: bounded here by scan synthesis time)",
                       "iterations": 4, "sequences": n_seq, "seqs_per_gpu": S, "n_raw_mean": pts_v / (n_seq * K),
                       "n_down_mean": float(sum(s[1] for s in st_v)) / (S * K), "effct_feat_mean": float(sum(s[2] for s in st_v)) / (S * K),
                       "map_points_per_sequence": int(np.mean([len(w["map_pts"]) for w in works])),
                       "l2": "not flushed: the concurrent sequences' maps (S x ~70 MB of buckets + tables) exceed the 126 MB L2 for S >= 2",
                       "timing": "one CUDA event before dlt_lio_replay_sequences to the last sequence's end event (max over sequences and ranks)", "worker_threads_per_gpu": n_thr,
                       "parallelism": "independent sequences, no data-path collective"},
            "e2e": {"value": pts_e / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(pts_e / (world * K) * 48), "d2h_bytes_per_step": int(S * 18064),
                    "ms_per_step": ms_e / K, "api": "dlt_lio_replay_sequences with pinned host buffers: dlt_lio_prefetch_scan(next) + dlt_lio_process_scan per scan"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": None, "cpu_baseline": None,
        }
        # whole-job HBM figure (SURVEY.md 8d formula for the algorithmic bytes of one scan, 2 match passes / 3 iterations as measured
        # on these scans) -- the configuration whose working set (S maps of ~70 MB) exceeds the 126 MB L2
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
            peak = float(peaks.get("hbm_gbs", 6650.0))
            nraw, ndn = line["config"]["n_raw_mean"], line["config"]["n_down_mean"]
            b_scan = nraw * 48 + nraw * 16 + ndn * 16 + 2 * ndn * 116 + 3 * ndn * 32 + 3 * 92 * 8
            ach = b_scan * line["scans_per_s"] / world / 1e9
            line["roofline"] = {"bound": "hbm", "kernel": "whole job, per GPU", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                                "algorithmic_bytes_per_scan": b_scan, "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                                "note": "many concurrent sequences: the maps no longer fit the L2, yet the job moves only this fraction of what HBM could deliver -- the update is a chain "
                                        "of dependent gathers with a host decision per iteration, bounded by instruction issue and launch rate, not by bytes (DESIGN.md section 5)"}
        except Exception as e:
            line["roofline"] = {"error": repr(e)}
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
