#!/usr/bin/env python
"""Measured parity series: the device pipeline and the oracle driven over the same raw scans, per scan

  * chained    both chains carry their own state and map from scan 0 on (what tests/test_pipeline_parity.py bounds);
  * re-synced  the device gets the oracle's state and map before every scan (identical inputs per scan; the north star's
               "pose within 1e-5 of the reference").

Per scan: pose error, feats_down_size on both sides, VoxelGrid voxels that differ, effct_feat_num, added points, map
points that differ beyond 1 mm.  Writes one JSON file; the chained-replay tolerances in the tests cite it.

    python tools/parity_series.py --scans 100 --out profiles/r02_parity_series.json            # GPU box
    python tools/parity_series.py --emu --scans 6                                              # kernel-logic emulator (CPU)
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scans", type=int, default=100)
    ap.add_argument("--emu", action="store_true")
    ap.add_argument("--workload", default="mid", choices=["small", "mid", "c2", "firstscan"])
    ap.add_argument("--out", default=None)
    args = ap.parse_args()

    import helpers
    import oracle_binding
    from daliti_b200 import synth
    from daliti_b200.binding import load_library
    from parity_tools import ParityRun

    lib = load_library(os.path.join(ROOT, "tests", "emu", "libdaliti_emu.so")) if args.emu else load_library()
    orc = oracle_binding.load()
    scans = None
    if args.workload == "c2":
        import bench

        work = bench.build_workload(0, args.scans, "c2")
        seq, map_pts, scans, name, thr = work["seq"], work["map_pts"], work["scans"], work["name"], 30
    elif args.workload == "firstscan":  # no prebuilt map: the first scan builds it (laserMapping.cpp:780-793), the chains start on a sparse map
        seq = helpers.small_sequence(seed=11, half=50.0, beams=32, azimuths=1024, n_boxes=20, speed=2.0, yaw_rate=0.2)
        map_pts, name, thr = None, "32x1024 scans, 100 m box world, map built from the first scan", 5
    elif args.workload == "mid":
        seq = helpers.small_sequence(seed=71, half=60.0, beams=32, azimuths=1024, n_boxes=30, speed=2.0, yaw_rate=0.2)
        map_pts, name, thr = synth.sample_map(seq.scene, seed=71), "32x1024 scans, 120 m box world", 5
    else:
        seq = helpers.small_sequence(seed=71, half=25.0, beams=16, azimuths=240, n_boxes=8, speed=2.0, yaw_rate=0.2)
        map_pts, name, thr = synth.sample_map(seq.scene, seed=71), "16x240 scans, 50 m box world", 5
    out = {"workload": name, "map_points": int(len(map_pts)) if map_pts is not None else 0, "scans": args.scans, "backend": "emulator" if args.emu else "B200"}
    for mode in ("chained", "resynced"):
        run = ParityRun(lib, orc, seq, map_pts, threads=min(8, os.cpu_count() or 1), chained=(mode == "chained"), identical=(mode != "chained"),
                        featptsThreshold=thr, max_scan_points=1 << 18 if args.workload == "c2" else 1 << 16)
        for k in range(args.scans):
            run.step(k, scans[k] if scans is not None else None)
        rec = run.summary()
        pe = np.array([r["pose_rel_err"] for r in run.series])
        rec["pose_rel_err_p50"], rec["pose_rel_err_p99"], rec["pose_rel_err_max"] = float(np.median(pe)), float(np.percentile(pe, 99)), float(pe.max())
        rec["n_down_abs_diff_max"] = int(max(abs(r["n_down_dev"] - r["n_down_orc"]) for r in run.series))
        out[mode] = {"summary": rec, "series": run.series}
        run.close()
        print(mode, json.dumps({k: v for k, v in rec.items()}), file=sys.stderr)
    txt = json.dumps(out)
    if args.out:
        with open(args.out, "w") as f:
            f.write(txt + "\n")
    else:
        print(txt)


if __name__ == "__main__":
    main()
