#!/usr/bin/env python
"""Stage times inside the persistent loop kernel (k_iekf_loop) on the C2 workload: the solver block's globaltimer stamps per
iteration (chunk work = match pass + residual pass of its own chunk, wait for the other blocks, final reduce, solve /
control step, grid barrier) and the SM-clock stamps inside iekf_step_block.  Development tool."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import bench
    from daliti_b200.lio import LaserMapping

    n, warm = 25, 5
    mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    work = bench.build_workload(0, n, "c2")
    seq, scans = work["seq"], work["scans"]
    lm = LaserMapping(dev=dict(device=0, max_scan_points=1 << 18, max_map_points=max(1 << 22, 2 * len(work["map_pts"]))), featptsThreshold=30,
                      device_loop=mode)
    s0, mean_acc, last_imu = bench.initial_state(seq)
    lm.force_imu_ready(mean_acc, last_imu)
    lm.set_state(s0)
    lm.device.map_build(work["map_pts"])
    devs = [torch.from_numpy(np.ascontiguousarray(p)).cuda() for p, _, _ in scans]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
    acc, cnt = {}, {}
    for k in range(n):
        pts, t_beg, imu = scans[k]
        lm.on_lidar_msg()
        flush.zero_()
        torch.cuda.synchronize()
        o = lm.process_scan_dev(devs[k].data_ptr(), len(pts), t_beg, t_beg + float(pts[-1, 6]), imu)
        if k < warm:
            continue
        ck = np.zeros((4, 16), np.int64)
        lm.lib.dlt_get_iekf_clocks(lm.device.h, ck.ctypes.data_as(C.c_void_p), C.c_int(4))
        its = lm.iters()
        for it in range(o.n_iters):
            key = (it, int(its[it].did_match))
            g = ck[it, 9:15]
            row = np.concatenate([np.diff(g) * 1e-3, [(g[-1] - g[0]) * 1e-3], (ck[it, 1:9] - ck[it, 0:8]) / 1.9e3])
            acc[key] = acc.get(key, 0) + row
            cnt[key] = cnt.get(key, 0) + 1
    print("per iteration (us): own chunks | wait others | final reduce | step | barrier | total   ||  step stages (us at 1.9 GHz): load, window, Q+boxminus, (sync), GJ, K1+sol, boxplus, store+ctl, blend")
    for key in sorted(acc):
        r = acc[key] / cnt[key]
        print(f"  iter {key[0]} match={key[1]} (n={cnt[key]}): " + " | ".join(f"{v:6.1f}" for v in r[:6]) + "  ||  " + " ".join(f"{v:5.2f}" for v in r[6:]))
    lm.close()


if __name__ == "__main__":
    main()
