"""Shared helpers for the parity tests (sequence setup, state packing)."""
from __future__ import annotations

import numpy as np

from daliti_b200 import synth


def state612(traj: synth.Trajectory, t: float, R_L_I=None, T_L_I=None) -> np.ndarray:
    """StatesGroup flat layout: rot_end[9] pos_end[3] R_L_I[9] T_L_I[3] vel[3] bg[3] ba[3] g[3] cov[576]."""
    s = np.zeros(36 + 576)
    s[0:24] = traj.pose24(t, R_L_I, T_L_I)
    s[24:27] = traj.vel(t)
    s[33:36] = [0, 0, -9.801]
    s[36:] = np.eye(24).ravel()
    return s


def small_sequence(seed=0, half=30.0, beams=16, azimuths=360, speed=1.0, yaw_rate=0.1, n_boxes=10, tunnel=False):
    if tunnel:
        scene = synth.make_tunnel(length=200.0)
        traj = synth.Trajectory(speed=speed, yaw_rate=0.0, x0=-20.0, z0=1.2)
        spec = synth.ScanSpec(beams, azimuths, (-22.5, 22.5), max_range=40.0)  # end caps out of range
    else:
        scene = synth.make_box_world(half=half, n_boxes=n_boxes, seed=seed, keep_clear=4.0)
        traj = synth.Trajectory(speed=speed, yaw_rate=yaw_rate, z0=1.5)
        spec = synth.ScanSpec(beams, azimuths, (-15.0, 15.0), max_range=80.0)
    return synth.Sequence(scene, traj, spec, seed=seed)


def start_oracle_lio(oracle, seq: synth.Sequence, map_pts, map_kind, **cfg_kw):
    """An oracle pipeline with IMU initialisation skipped, the state at ground truth and a prebuilt map."""
    import oracle_binding as ob

    lio = oracle.new_lio(ob.default_lio_config(**cfg_kw), map_kind)
    lio.force_imu_ready([0.0, 0.0, synth.G], np.concatenate([[seq.t_start - 0.005], [0, 0, synth.G], [0, 0, seq.traj.yaw_rate]]))
    lio.set_state(state612(seq.traj, seq.t_start))
    if map_pts is not None:
        lio.map().build(map_pts)
    return lio
