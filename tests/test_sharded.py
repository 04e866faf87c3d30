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


def _run_lio(lib, rank, world, n_scans=3, peers=False, device_loop=-1, device=0, native_nccl=False):
    from daliti_b200.lio import LaserMapping

    seq = helpers.small_sequence(seed=22, half=30.0, beams=16, azimuths=240, n_boxes=8)
    map_pts = synth.sample_map(seq.scene, seed=22)
    lm = LaserMapping(lib, dev=dict(max_scan_points=8192, max_map_points=1 << 17, shard_rank=rank, shard_count=world, shard_tile_shift=3, device=device),
                      featptsThreshold=5, device_loop=device_loop)
    lm.force_imu_ready([0.0, 0.0, synth.G], np.concatenate([[seq.t_start - 0.005], [0, 0, synth.G], [0, 0, seq.traj.yaw_rate]]))
    lm.set_state(helpers.state612(seq.traj, seq.t_start))
    lm.device.map_build(map_pts)
    if world > 1 and peers:  # sums over the ranks inside the kernels, through the peer mailboxes; no callback at all
        from daliti_b200.sharded import attach_peers

        assert attach_peers(lm)
    elif world > 1 and native_nccl:  # include/daliti_b200_nccl.h: ncclAllReduce issued from C on the handle's stream
        from daliti_b200.sharded import attach_native_nccl

        attach_native_nccl(lm, device)
    elif world > 1:
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


def _lio_worker_all(rank, world, port, q, peers=False, device_loop=-1, gpu=False, native_nccl=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    from daliti_b200.binding import load_library

    # (gloo only carries the 128-byte mailbox blobs, once; on GPUs the data path is NVLink peer memory)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = load_library() if gpu else load_library(os.path.join(ROOT, "tests", "emu", "libdaliti_emu.so"))
    states = _run_lio(lib, rank, world, peers=peers, device_loop=device_loop, device=rank if gpu else 0, native_nccl=native_nccl)
    q.put((rank, states))
    dist.destroy_process_group()


def test_sharded_per_scan_update_four_processes_peers(emu_lib):
    """four ranks on the peer mailboxes: every rank posts to three peers and adds four slots in rank order"""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 35500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_lio_worker_all, args=(r, 4, port, q, True, -1)) for r in range(4)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in range(4))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ref = _run_lio(emu_lib, 0, 1)
    _check_sharded_against_unsharded(got, ref, bitwise_ranks=True)


@pytest.mark.parametrize("mode", ["callback", "peers", "peers_host_loop", "peers_device_finish"])
def test_sharded_per_scan_update_two_processes(emu_lib, mode):
    """the full per-scan update on a 2-way sharded map -- sum of the normal equations over the ranks every iteration inside
    the device-resident loop, replicated 24-state solve, map_incremental with the owners' decisions exchanged -- follows the
    unsharded update: effective counts exact, states to fp64 rounding, and the union of the owned tiles holds exactly
    the unsharded map after every scan.  callback: the sums go through the reduce callback (gloo here, NCCL on GPUs);
    peers: through the peer mailboxes from inside k_residual / k_incr_push (shared memory between the two emulator
    processes here, NVLink peer memory on GPUs), with the solve step fused behind the exchange; peers_host_loop: the same
    with one dlt_measure round trip per iteration; peers_device_finish: device_loop = 1 -- zeta blend and map_incremental
    (classification, k_incr_push / k_incr_pull, insert) queued on the device behind the loop, one synchronisation per scan."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 500) * 4 + {"callback": 0, "peers": 1, "peers_host_loop": 2, "peers_device_finish": 3}[mode]  # (x4: parallel workers have consecutive pids)
    peers = mode != "callback"
    loop = {"peers_host_loop": 0, "peers_device_finish": 1}.get(mode, -1)
    procs = [ctx.Process(target=_lio_worker_all, args=(r, 2, port, q, peers, loop)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ref = _run_lio(emu_lib, 0, 1)
    _check_sharded_against_unsharded(got, ref, bitwise_ranks=peers)


def _check_sharded_against_unsharded(got, ref, bitwise_ranks):
    for k in range(len(ref)):
        st, eff, n_it, added, pts = ref[k]
        for r in range(len(got)):
            st_r, eff_r, n_it_r, added_r, _ = got[r][k]
            assert n_it_r == n_it and eff_r == eff, (k, r)
            assert added_r == added, (k, r, added_r, added)
            np.testing.assert_allclose(st_r, st, rtol=1e-8, atol=1e-8, err_msg=f"scan {k} rank {r}")
        if bitwise_ranks:  # the mailboxes are added up in rank order on every rank: the replicated solves see the same bits
            for r in range(1, len(got)):
                assert np.array_equal(got[0][k][0], got[r][k][0]), (k, r)
        # the sharded run's poses differ from the unsharded ones by the summation order of the partial normal equations
        # (~1e-10), so inserted points may differ in their last float32 bit: match point for point within 1e-5 m
        from scipy.spatial import cKDTree

        owned = np.concatenate([got[r][k][4] for r in range(len(got))], 0)
        assert len(owned) == len(pts), (k, len(owned), len(pts))
        d, idx = cKDTree(pts[:, :3]).query(owned[:, :3])
        assert d.max() < 1e-5, (k, d.max())
        assert len(np.unique(idx)) == len(pts)  # a bijection: no tile is held twice, none is missing
        np.testing.assert_allclose(owned[:, 3], pts[idx, 3], rtol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("device_loop", [0, 2])
def test_sharded_per_scan_update_two_gpus_native_nccl(gpu_lib, device_loop):
    """two B200s, one process per GPU, the library's own NCCL transport (daliti_b200_nccl.h): ncclAllReduce of the 158 doubles
    issued from C between the residual pass and the solve (and of the per-point map_incremental decisions); torch.distributed
    only carries the 128-byte unique id once."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    if not gpu_lib.dlt_nccl_available():
        pytest.skip("libnccl.so.2 not loadable")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 36500 + (os.getpid() % 500) * 4 + device_loop
    procs = [ctx.Process(target=_lio_worker_all, args=(r, 2, port, q, False, device_loop, True, True)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ref = _run_lio(gpu_lib, 0, 1)
    _check_sharded_against_unsharded(got, ref, bitwise_ranks=False)


def test_native_nccl_api_without_nccl_or_gpu(dev):
    """the transport is optional: argument errors come back as codes, and a missing libnccl is reported, not fatal"""
    import ctypes as C

    lib, is_gpu = dev
    lib.dlt_nccl_last_error.restype = C.c_char_p
    assert lib.dlt_nccl_unique_id(None) != 0
    comm = C.c_void_p()
    assert lib.dlt_nccl_create(None, 0, 1, 0, C.byref(comm)) != 0
    assert lib.dlt_nccl_allreduce(None, None, 4) != 0
    assert lib.dlt_nccl_available() in (0, 1)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["peers", "peers_host_loop"])
def test_sharded_per_scan_update_two_gpus_peer_memory(gpu_lib, mode):
    """the same on two B200s, one process per GPU: mailboxes in HBM mapped into the peer process by CUDA IPC, the partial
    normal equations and the map_incremental decisions stored over NVLink from inside the kernels -- no NCCL, no callback."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 500) * 4 + (1 if mode == "peers" else 2)
    procs = [ctx.Process(target=_lio_worker_all, args=(r, 2, port, q, True, 0 if mode == "peers_host_loop" else -1, True)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ref = _run_lio(gpu_lib, 0, 1)
    _check_sharded_against_unsharded(got, ref, bitwise_ranks=True)


def test_peer_exchange_times_out_instead_of_hanging(dev):
    """a rank whose peer never takes part gives up after 5 s and reports it; the device is not left spinning"""
    import time

    from daliti_b200.binding import DltError

    lib, is_gpu = dev
    seq, map_pts, pts = scene_scan()
    down = downsample_body(lib, pts)
    hs = [ScanToMap(lib, max_scan_points=8192, max_map_points=1 << 17, shard_rank=r, shard_count=2, shard_tile_shift=3) for r in range(2)]
    blobs = [h.peer_export() for h in hs]
    for h in hs:
        h.map_build(map_pts)
        h.peer_attach(blobs)  # same process: the mailboxes are addressed directly
    with pytest.raises(DltError, match="attached"):
        hs[0].peer_attach(blobs)
    hs[0].scan_set_down(down)
    t0 = time.time()
    with pytest.raises(DltError, match="timed out"):
        hs[0].measure(seq.traj.pose24(0.1), True)  # rank 1 never calls
    assert 4.0 < time.time() - t0 < 30.0
    for h in hs:
        h.close()


def test_peer_attach_argument_errors(dev):
    from daliti_b200.binding import DltError

    lib, is_gpu = dev
    one = ScanToMap(lib, max_scan_points=4096, max_map_points=4096)
    with pytest.raises(DltError):
        one.peer_export()  # not a sharded handle
    one.close()
    a = ScanToMap(lib, max_scan_points=4096, max_map_points=4096, shard_rank=0, shard_count=2, shard_tile_shift=3)
    b = ScanToMap(lib, max_scan_points=8192, max_map_points=4096, shard_rank=1, shard_count=2, shard_tile_shift=3)
    with pytest.raises(DltError, match="before dlt_peer_export"):
        a.peer_attach([bytes(128), bytes(128)])
    ba, bb = a.peer_export(), b.peer_export()
    with pytest.raises(DltError, match="rank order"):
        a.peer_attach([bb, ba])
    with pytest.raises(DltError, match="size differs"):
        a.peer_attach([ba, bb])  # max_scan_points differ
    with pytest.raises(DltError, match="cannot map"):
        a.peer_attach([ba, bytes(128)])
    a.close()
    b.close()
