#!/usr/bin/env python
"""A short, deterministic replay of the bench workload for ncu: W warm-up scans + K scans of the
C2 configuration through dlt_lio_process_scan_dev.  (Numbers printed under a profiler are never
bench values.)"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scans", type=int, default=6)
    ap.add_argument("--workload", default="c2")
    args = ap.parse_args()
    import torch

    import bench
    from daliti_b200.lio import LaserMapping

    work = bench.build_workload(0, args.scans, args.workload)
    seq, scans = work["seq"], work["scans"]
    lm = LaserMapping(dev=dict(device=0, max_scan_points=1 << 18, max_map_points=max(1 << 22, 2 * len(work["map_pts"]))), featptsThreshold=30)
    s0, mean_acc, last_imu = bench.initial_state(seq)
    lm.force_imu_ready(mean_acc, last_imu)
    lm.set_state(s0)
    lm.device.map_build(work["map_pts"])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
    for k in range(args.scans):
        pts, t_beg, imu = scans[k]
        d = torch.from_numpy(np.ascontiguousarray(pts)).cuda()
        lm.on_lidar_msg()
        flush.zero_()
        torch.cuda.synchronize()
        o = lm.process_scan_dev(d.data_ptr(), len(pts), t_beg, t_beg + float(pts[-1, 6]), imu)
        print(k, o.n_raw, o.n_down, o.n_iters, o.added, f"{1e3 * o.t_total:.3f} ms")
    lm.close()


if __name__ == "__main__":
    main()
