#!/usr/bin/env python
"""A/B of the per-scan update latency between configurations of the host loop (device_loop = 0 host loop, 2 device loop +
host-driven blend/insert, 1 device loop + device-side finish, 0z host loop with the zero-copy result block) on the C2 workload: p50 of per-scan CUDA-event times with
the L2 flushed between scans, and the host wall clock.  Development tool."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import bench
    from daliti_b200.lio import LaserMapping

    n, warm = int(os.environ.get("AB_SCANS", "45")), 5
    work = bench.build_workload(0, n, "c2")
    seq, scans = work["seq"], work["scans"]
    devs = [torch.from_numpy(np.ascontiguousarray(p)).cuda() for p, _, _ in scans]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
    stream = torch.cuda.Stream()
    # first character: device_loop; then flags: g = CUDA graph with conditional nodes, n = NO zero-copy result block (host loop),
    # c = classic three kernels per iteration (DLT_LOOP_FUSED=0), p = the persistent cooperative loop kernel (DLT_LOOP_COOP=1);
    # default for 1 / 2: one fused launch per iteration
    modes = [a for a in sys.argv[1:]] or ["0", "0n", "1", "1p", "1c", "2"]
    for rep in range(2):
        for mtag in modes:
            mode = int(mtag[0])
            os.environ["DLT_LOOP_GRAPH"] = "1" if "g" in mtag[1:] else "0"
            os.environ["DLT_ZEROCOPY"] = "0" if "n" in mtag[1:] else "1"
            os.environ["DLT_LOOP_FUSED"] = "0" if "c" in mtag[1:] else "1"
            os.environ["DLT_LOOP_COOP"] = "1" if "p" in mtag[1:] else "0"
            os.environ["DLT_PDL"] = "0" if "s" in mtag[1:] else "1"  # s = plain stream serialisation (no programmatic dependent launch)
            os.environ["DLT_KNN_REUSE"] = "0" if "r" in mtag[1:] else "1"  # r = every match pass searches (no reuse of proven neighbour sets)
            async_insert = 0 if "i" in mtag[1:] else 1                     # i = map_incremental inside the call (synchronous)
            lm = LaserMapping(dev=dict(device=0, max_scan_points=1 << 18, max_map_points=max(1 << 22, 2 * len(work["map_pts"]))), featptsThreshold=30,
                              device_loop=mode, async_insert=async_insert)
            lm.collect_after_scan = False
            lm.device.set_stream(stream.cuda_stream)
            s0, mean_acc, last_imu = bench.initial_state(seq)
            lm.force_imu_ready(mean_acc, last_imu)
            lm.set_state(s0)
            lm.device.map_build(work["map_pts"])
            ev, host, stages = [], [], []
            l0 = lm.device.launch_count()
            with torch.cuda.stream(stream):
                for k in range(n):
                    pts, t_beg, imu = scans[k]
                    lm.on_lidar_msg()
                    flush.zero_()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(stream)
                    t0 = time.perf_counter()
                    o = lm.process_scan_dev(devs[k].data_ptr(), len(pts), t_beg, t_beg + float(pts[-1, 6]), imu)
                    host.append(1e3 * (time.perf_counter() - t0))
                    stages.append((o.n_iters, o.n_down, [round(1e3 * x, 3) for x in (o.t_deskew, o.t_voxel, o.t_iterate, o.t_insert, o.t_delete, o.t_total)]))
                    b.record(stream)
                    ev.append((a, b))
                torch.cuda.synchronize()
            ms = np.array([a.elapsed_time(b) for a, b in ev])[warm:]
            st = lm.get_state()
            print(f"device_loop={mtag}: p50 {np.median(ms):.4f} ms  mean {ms.mean():.4f}  host p50 {np.median(host[warm:]):.4f} ms   launches/scan "
                  f"{(lm.device.launch_count() - l0) / n:.1f}  final pos {st[9]:.6f} {st[10]:.6f} {st[11]:.6f}", flush=True)
            slow = [int(i) + warm for i in np.argsort(-ms)[:3] if ms[i] > 1.3 * np.median(ms)]
            for i in slow:
                print(f"    slow scan {i}: {ms[i - warm]:.4f} ms  host {host[i]:.4f} ms  iters/n_down/host stages {stages[i]}", flush=True)
            lm.close()


if __name__ == "__main__":
    main()
