"""Spatially sharded map (BASELINE config C4): every rank holds the tiles it owns plus a
halo, evaluates only the query points whose tile it owns, and the partial normal equations are summed.
  * in-process: 2 and 4 shard handles, sum of partials == unsharded result (E exact, fp64 rel 1e-12)
  * 2 processes over torch.distributed (gloo on CPU with the emulator library here; the same code path
    runs over NCCL on the GPU box in bench.py --workload c4)."""
import os
import sys

import numpy as np
import pytest

import helpers
from daliti_b200 import synth
from daliti_b200.binding import ScanToMap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def scene_scan():
    seq = helpers.small_sequence(seed=21, half=30.0, beams=16, azimuths=240, n_boxes=8)
    map_pts = synth.sample_map(seq.scene, seed=21)
    pts, t_beg, imu = seq.scan(0)
    return seq, map_pts, pts


def downsample_body(lib, pts):
    dm = ScanToMap(lib, max_scan_points=8192, max_map_points=4096)
    dm.scan_deskew(pts)
    n = dm.scan_downsample()
    down = dm.scan_get_down(n)
    dm.close()
    return down


@pytest.mark.parametrize("shards", [2, 4])
def test_sharded_partials_sum_to_unsharded(dev, shards):
    lib, is_gpu = dev
    seq, map_pts, pts = scene_scan()
    down = downsample_body(lib, pts)
    pose = seq.traj.pose24(0.1)
    full = ScanToMap(lib, max_scan_points=8192, max_map_points=1 << 17)
    full.map_build(map_pts)
    full.scan_set_down(down)
    m0 = full.measure(pose, True)
    m1 = full.measure(pose, False)
    HtH = np.zeros((12, 12))
    Htr = np.zeros(12)
    eff = 0
    res = 0.0
    live = 0
    HtH1 = np.zeros((12, 12))
    for r in range(shards):
        sh = ScanToMap(lib, max_scan_points=8192, max_map_points=1 << 17, shard_rank=r, shard_count=shards, shard_tile_shift=3)
        sh.map_build(map_pts)
        live += sh.map_valid_count()
        sh.scan_set_down(down)
        m = sh.measure(pose, True)
        HtH += m.HtH
        Htr += m.Htr
        eff += m.effct_feat_num
        res += m.total_residual
        HtH1 += sh.measure(pose, False).HtH
        sh.close()
    assert live >= len(map_pts)  # halos are replicated
    assert live < shards * len(map_pts)
    assert eff == m0.effct_feat_num
    np.testing.assert_allclose(HtH, m0.HtH, rtol=1e-12, atol=1e-12 * abs(m0.HtH).max())
    np.testing.assert_allclose(Htr, m0.Htr, rtol=1e-10, atol=1e-12 * abs(m0.Htr).max())
    np.testing.assert_allclose(res, m0.total_residual, rtol=1e-12)
    np.testing.assert_allclose(HtH1, m1.HtH, rtol=1e-12, atol=1e-12 * abs(m1.HtH).max())
    full.close()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    from daliti_b200.binding import load_library
    from daliti_b200.sharded import allreduce_measure

    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = load_library(os.path.join(ROOT, "tests", "emu", "libdaliti_emu.so"))
    seq, map_pts, pts = scene_scan()
    down = downsample_body(lib, pts)
    pose = seq.traj.pose24(0.1)
    sh = ScanToMap(lib, max_scan_points=8192, max_map_points=1 << 17, shard_rank=rank, shard_count=world, shard_tile_shift=3)
    sh.map_build(map_pts)
    sh.scan_set_down(down)
    out = allreduce_measure(sh, pose, True, device="cpu")
    if rank == 0:
        q.put((out["HtH"], out["Htr"], out["effct_feat_num"], out["total_residual"]))
    dist.destroy_process_group()


def test_two_process_allreduce_gloo(emu_lib):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    HtH, Htr, eff, res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    seq, map_pts, pts = scene_scan()
    down = downsample_body(emu_lib, pts)
    full = ScanToMap(emu_lib, max_scan_points=8192, max_map_points=1 << 17)
    full.map_build(map_pts)
    full.scan_set_down(down)
    m0 = full.measure(seq.traj.pose24(0.1), True)
    assert eff == m0.effct_feat_num
    np.testing.assert_allclose(HtH, m0.HtH, rtol=1e-12, atol=1e-12 * abs(m0.HtH).max())
    np.testing.assert_allclose(Htr, m0.Htr, rtol=1e-10, atol=1e-12 * abs(m0.Htr).max())
    np.testing.assert_allclose(res, m0.total_residual, rtol=1e-12)
    full.close()


def _lio_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    from daliti_b200.binding import load_library
    from daliti_b200.lio import LaserMapping

    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = load_library(os.path.join(ROOT, "tests", "emu", "libdaliti_emu.so"))
    states = _run_lio(lib, rank, world)
    if rank == 0:
        q.put(states)
    dist.destroy_process_group()


def _tile_owner(xyz, shards, ds=0.5, cell_shift=1, tile_shift=3):
    """numpy restatement of dlt::tile_owner (dlt_common.cuh) for the cell of each point"""
    v = np.floor(xyz.astype(np.float32) / np.float32(ds)).astype(np.int64)
    t = (v >> cell_shift) >> tile_shift
    m = np.uint64(0x1FFFFF)
    k = ((t[:, 0].astype(np.uint64) & m) << np.uint64(42)) | ((t[:, 1].astype(np.uint64) & m) << np.uint64(21)) | (t[:, 2].astype(np.uint64) & m)
    with np.errstate(over="ignore"):
        k ^= k >> np.uint64(33)
        k *= np.uint64(0xFF51AFD7ED558CCD)
        k ^= k >> np.uint64(33)
        k *= np.uint64(0xC4CEB9FE1A85EC53)
        k ^= k >> np.uint64(33)
    return ((k & np.uint64(0xFFFFFFFF)) % np.uint64(shards)).astype(np.int64)


def _run_lio(lib, rank, world, n_scans=3):
    from daliti_b200.lio import LaserMapping

    seq = helpers.small_sequence(seed=22, half=30.0, beams=16, azimuths=240, n_boxes=8)
    map_pts = synth.sample_map(seq.scene, seed=22)
    lm = LaserMapping(lib, dev=dict(max_scan_points=8192, max_map_points=1 << 17, shard_rank=rank, shard_count=world, shard_tile_shift=3),
                      featptsThreshold=5)
    lm.force_imu_ready([0.0, 0.0, synth.G], np.concatenate([[seq.t_start - 0.005], [0, 0, synth.G], [0, 0, seq.traj.yaw_rate]]))
    lm.set_state(helpers.state612(seq.traj, seq.t_start))
    lm.device.map_build(map_pts)
    if world > 1:
        lm.set_allreduce("cpu")
    states = []
    for k in range(n_scans):
        pts, t_beg, imu = seq.scan(k)
        lm.on_lidar_msg()
        o = lm.process_scan(pts, t_beg, imu)
        m = lm.device.map_export()
        if world > 1:  # only the points of the tiles this rank owns (halos are replicas)
            m = m[_tile_owner(m[:, :3], world) == rank]
        states.append((lm.get_state()[:36].copy(), [it.effct_feat_num for it in lm.iters()], o.n_iters, o.added, m.astype(np.float64)))
    lm.close()
    return states


def _lio_worker_all(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    from daliti_b200.binding import load_library

    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = load_library(os.path.join(ROOT, "tests", "emu", "libdaliti_emu.so"))
    states = _run_lio(lib, rank, world)
    q.put((rank, states))
    dist.destroy_process_group()


def test_sharded_per_scan_update_two_processes(emu_lib):
    """the full per-scan update on a 2-way sharded map -- all-reduce of the normal equations every iteration inside the
    device-resident loop, replicated 24-state solve, map_incremental with the owners' decisions exchanged -- follows the
    unsharded update: effective counts exact, states to fp64 rounding, and the union of the owned tiles holds exactly
    the unsharded map after every scan."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_lio_worker_all, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ref = _run_lio(emu_lib, 0, 1)
    for k in range(len(ref)):
        st, eff, n_it, added, pts = ref[k]
        for r in range(2):
            st_r, eff_r, n_it_r, added_r, _ = got[r][k]
            assert n_it_r == n_it and eff_r == eff, (k, r)
            assert added_r == added, (k, r, added_r, added)
            np.testing.assert_allclose(st_r, st, rtol=1e-8, atol=1e-8, err_msg=f"scan {k} rank {r}")
        # the sharded run's poses differ from the unsharded ones by the summation order of the partial normal equations
        # (~1e-10), so inserted points may differ in their last float32 bit: match point for point within 1e-5 m
        from scipy.spatial import cKDTree

        owned = np.concatenate([got[0][k][4], got[1][k][4]], 0)
        assert len(owned) == len(pts), (k, len(owned), len(pts))
        d, idx = cKDTree(pts[:, :3]).query(owned[:, :3])
        assert d.max() < 1e-5, (k, d.max())
        assert len(np.unique(idx)) == len(pts)  # a bijection: no tile is held twice, none is missing
        np.testing.assert_allclose(owned[:, 3], pts[idx, 3], rtol=1e-5)
