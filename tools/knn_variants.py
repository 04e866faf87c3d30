#!/usr/bin/env python
"""Times k_knn / k_residual of several builds of the library (gpurun_scratch/*.so; `make -C daliti_b200/csrc variants` builds the
opt-in ones, e.g. DLT_KNN8_PRUNE) on the C2 workload.
Development tool for choosing launch bounds / load batching; not a bench value."""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import bench
    from daliti_b200.binding import ScanToMap, load_library

    work = bench.build_workload(0, 2, "c2")
    seq = work["seq"]
    pts, t_beg, imu = work["scans"][1]
    pose = seq.traj.pose24(t_beg + 0.1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
    libs = sorted(glob.glob(os.path.join(ROOT, "gpurun_scratch", "*.so"))) + [os.path.join(ROOT, "daliti_b200", "lib", "libdaliti_b200.so")]
    down = None
    for path in libs:
        lib = load_library(path)
        dm = ScanToMap(lib, max_scan_points=1 << 18, max_map_points=1 << 22)
        dm.map_build(work["map_pts"])
        if down is None:
            dm.scan_deskew(pts)
            n = dm.scan_downsample()
            down = dm.scan_get_down(n)
        dm.scan_set_down(down)
        dm.set_profiling(True)
        for _ in range(3):
            dm.measure(pose, True)
        dm.get_profile(reset=True)
        for _ in range(20):
            flush.zero_()
            torch.cuda.synchronize()
            m = dm.measure(pose, True)
        prof = dm.get_profile(reset=True)
        print(f"{os.path.basename(path):24s} n_down={len(down)} effct={m.effct_feat_num} knn {1e3*prof['knn'][0]/prof['knn'][1]:7.1f} us  knn8 {1e3*prof['knn8'][0]/max(1,prof['knn8'][1]):7.1f} us  residual {1e3*prof['residual'][0]/prof['residual'][1]:7.1f} us", flush=True)
        dm.close()


if __name__ == "__main__":
    main()
