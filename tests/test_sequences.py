"""dlt_lio_replay_sequences: many independent sequences replayed natively (worker threads x coroutines that switch where the
library waits for the device) must give exactly what the same calls give one by one."""
import numpy as np
import pytest

import helpers
from daliti_b200 import synth
from daliti_b200.lio import LaserMapping, replay_sequences


def _make(lib, seed, device_loop):
    seq = helpers.small_sequence(seed=seed, half=30.0, beams=16, azimuths=240, n_boxes=8)
    map_pts = synth.sample_map(seq.scene, seed=seed)
    lm = LaserMapping(lib, dev=dict(max_scan_points=8192, max_map_points=1 << 17), featptsThreshold=5, device_loop=device_loop)
    lm.force_imu_ready([0.0, 0.0, synth.G], np.concatenate([[seq.t_start - 0.005], [0, 0, synth.G], [0, 0, seq.traj.yaw_rate]]))
    lm.set_state(helpers.state612(seq.traj, seq.t_start))
    lm.device.map_build(map_pts)
    return seq, lm


def _final(lm):
    return lm.get_state().copy(), set(map(tuple, lm.device.map_export().tolist()))


def _check(lib, n_seq, n_scans, n_threads, device_loop=0):
    one_by_one, outs_ref = [], []
    for s in range(n_seq):
        seq, lm = _make(lib, 50 + s, device_loop)
        outs = []
        for k in range(n_scans):
            pts, t_beg, imu = seq.scan(k)
            lm.on_lidar_msg()
            o = lm.process_scan(pts, t_beg, imu)
            outs.append((o.n_raw, o.n_down, o.n_iters, o.ekf_stop, o.did_update))
        one_by_one.append(_final(lm))
        outs_ref.append(outs)
        lm.close()
    made = [_make(lib, 50 + s, device_loop) for s in range(n_seq)]
    scans = [[seq.scan(k) for k in range(n_scans)] for seq, _ in made]
    outs = replay_sequences([lm for _, lm in made], scans, n_threads=n_threads)
    for s, (_, lm) in enumerate(made):
        st, mp = _final(lm)
        np.testing.assert_array_equal(st, one_by_one[s][0])
        assert mp == one_by_one[s][1]
        assert [(o.n_raw, o.n_down, o.n_iters, o.ekf_stop, o.did_update) for o in outs[s]] == outs_ref[s]
        assert lm.flags()["lidar_cnt"] == n_scans
        lm.close()


def test_replay_sequences_equals_one_by_one(emu_lib):
    """one worker thread, three sequences as coroutines that switch at every device wait (the kernel-logic emulator itself is
    single-threaded; several worker threads run in the GPU test)"""
    _check(emu_lib, 3, 3, 1)


def test_replay_sequences_argument_errors(emu_lib):
    import ctypes as C

    from daliti_b200.lio import LioSeq
    assert emu_lib.dlt_lio_replay_sequences(None, C.c_int(0), C.c_int(1)) == 0
    assert emu_lib.dlt_lio_replay_sequences(None, C.c_int(2), C.c_int(1)) != 0
    seqs = (LioSeq * 1)()
    assert emu_lib.dlt_lio_replay_sequences(seqs, C.c_int(1), C.c_int(1)) != 0  # null handle


@pytest.mark.gpu
@pytest.mark.parametrize("device_loop", [0, 1])
def test_replay_sequences_equals_one_by_one_gpu(gpu_lib, device_loop):
    _check(gpu_lib, 6, 5, 2, device_loop)
