#!/usr/bin/env python
"""Prints the measured-results table of DESIGN.md section 5 from the bench lines committed under profiles/."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def load(name):
    path = os.path.join(P, name)
    return json.load(open(path)) if os.path.exists(path) else None


def main(tag="v5"):
    rows = ["| workload | p50 per scan (HBM-resident) | points/s | e2e (host buffers) | CPU reference, same box |", "|---|---|---|---|---|"]
    ref = load(f"r01_bench_ref_{tag}.json")
    for key, label in (("c2", "C2"), ("c1", "C1"), ("c3", "C3"), ("c4_n1", "C4, 1 GPU (55.6 M-pt map, unsharded)")):
        d = load(f"r01_bench_{key}_{tag}.json")
        if not d:
            continue
        e = d["e2e"]
        e2e = f"{e['ms_p50']:.3f} ms p50, {e['value'] / 1e6:.0f} M pts/s" if "ms_p50" in e else f"{e['ms_per_step']:.3f} ms / scan, {e['value'] / 1e6:.0f} M pts/s"
        cpu = "—"
        if key == "c2" and ref:
            pr = ref["config"]["threads_probe_points_per_s"]
            cpu = (f"{ref['ms_p50']:.1f} ms / scan, {ref['value'] / 1e6:.2f} M pts/s on {ref['cpu_baseline']['cores']} threads "
                   f"(as shipped, 1 thread: {pr['1'] / 1e6:.2f} M pts/s)")
        if key == "c3":
            cpu = f"— ({d['config']['deleted_total']} points box-deleted, {d['config']['degenerate_scans']} / {d['steps']} scans flagged degenerate)"
        rows.append(f"| {label} | {d['ms_p50']:.3f} ms | {d['value'] / 1e6:.0f} M | {e2e} | {cpu} |")
    for name, label in (("r01_bench_c4_n2_peer_v6.json", "C4, 2 GPUs (sharded, sums inside `k_residual` over NVLink peer mailboxes)"),
                        ("r01_bench_c4_n2_nccl_v6.json", "C4, 2 GPUs (sharded, NCCL all-reduce between `k_residual` and `k_iekf_step`; same run)"),
                        ("r01_bench_c4_n4_peer_v6.json", "C4, 4 GPUs (sharded, peer mailboxes)")):
        d = load(name)
        if d:
            p50 = f"{d['ms_p50']:.3f} ms" + (f" (mean {d['ms_per_step']:.3f})" if d["ms_per_step"] > 1.1 * d["ms_p50"] else "")
            rows.append(f"| {label} | {p50} | {d['value'] / 1e6:.0f} M | {d['e2e']['ms_per_step']:.3f} ms / scan | — |")
    for name, label in ((f"r01_bench_c5_s1_{tag}.json", "C5, 1 sequence on one GPU"), (f"r01_bench_c5_s16_{tag}.json", "C5, 16 concurrent sequences on one GPU"), (f"r01_bench_c5_s64_{tag}.json", "C5, 64 concurrent sequences on one GPU"),
                        ("r01_bench_c5_n2_v4.json", "C5, 2 GPUs x 8 sequences (device-resident loop)")):
        d = load(name)
        if d:
            rows.append(f"| {label} | {d['scans_per_s'] / 1e3:.1f} k scans/s aggregate | {d['value'] / 1e6:.0f} M | {d['e2e']['value'] / 1e6:.0f} M pts/s | — |")
    for name, label in (("r01_bench_c2_n2_v5.json", "C2, 2 GPUs (one sequence each)"), ("r01_bench_c2_n8_v4.json", "C2, 8 GPUs (one sequence each; measured before the last single-GPU trims)")):
        d = load(name)
        if d:
            rows.append(f"| {label} | {d['ms_p50']:.3f} ms | {d['value'] / 1e6:.0f} M | {d['e2e']['value'] / 1e6:.0f} M pts/s | — |")
    print("\n".join(rows))


if __name__ == "__main__":
    main(*sys.argv[1:])
