#!/usr/bin/env python
"""Device timeline of one per-scan update (CUDA events around every kernel group, on the launching streams) next to
the host-side stage clocks: where the gaps between kernels are.  Development tool; profiling events add ~1 us each."""
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

    n = int(os.environ.get("TL_SCANS", "8"))
    show = [int(x) for x in os.environ.get("TL_PRINT", f"{n - 2},{n - 1}").split(",")]
    wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
    work = bench.build_workload(0, n, wl)
    seq, scans = work["seq"], work["scans"]
    lm = LaserMapping(dev=dict(device=0, max_scan_points=1 << 18, max_map_points=max(1 << 22, 2 * len(work["map_pts"]))), featptsThreshold=30, **work["lm_kwargs"])
    s0, mean_acc, last_imu = bench.initial_state(seq)
    lm.force_imu_ready(mean_acc, last_imu)
    lm.set_state(s0)
    lm.device.map_build(work["map_pts"])
    devs = [torch.from_numpy(np.ascontiguousarray(p)).cuda() for p, _, _ in scans]
    for k in range(n):
        pts, t_beg, imu = scans[k]
        lm.on_lidar_msg()
        lm.device.set_profiling(k in show)  # (event pairs around every kernel group: only where the timeline is printed)
        lm.device.get_profile(reset=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        o = lm.process_scan_dev(devs[k].data_ptr(), len(pts), t_beg, t_beg + float(pts[-1, 6]), imu)
        host_ms = 1e3 * (time.perf_counter() - t0)
        tl = lm.device.get_timeline()
        if k in show:
            print(f"--- scan {k}: n_raw {o.n_raw} n_down {o.n_down} iters {o.n_iters} added {o.added} deleted {o.deleted} unresolved {[it.effct_feat_num for it in lm.iters()]}")
            print(f"--- scan {k}: host {host_ms:.3f} ms; stages deskew {1e3*o.t_deskew:.3f} voxel {1e3*o.t_voxel:.3f} iterate {1e3*o.t_iterate:.3f} insert {1e3*o.t_insert:.3f}")
            prev_end = 0.0
            busy = 0.0
            for name, a, b in tl:
                if name == "knn8":
                    continue  # nested inside knn
                gap = a - prev_end
                print(f"  {name:12s} start {1e3*a:8.1f} us  dur {1e3*(b-a):7.1f} us  gap before {1e3*gap:7.1f} us")
                if name != "eigen6":
                    prev_end = b
                    busy += b - a
            print(f"  device span {1e3*prev_end:.1f} us, busy {1e3*busy:.1f} us")
            import ctypes as C
            ck = np.zeros((4, 16), np.int64)
            lm.lib.dlt_get_iekf_clocks(lm.device.h, ck.ctypes.data_as(C.c_void_p), C.c_int(4))
            for it in range(4):
                d = ck[it, 1:9] - ck[it, 0:8]
                print("  k_iekf_step iter", it, "stage cycles:", d.tolist(), "total", int(ck[it, 8] - ck[it, 0]))
                print("     gj_warp: entry->", (ck[it, 9] - ck[it, 3]), "steps (pivot, elim):", (ck[it, 11] - ck[it, 10], ck[it, 12] - ck[it, 11]),
                      (ck[it, 13] - ck[it, 12], ck[it, 14] - ck[it, 13]), "exit->stamp4", ck[it, 4] - ck[it, 15] if ck[it, 15] else None)
    lm.close()


if __name__ == "__main__":
    main()
