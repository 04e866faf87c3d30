"""Deterministic synthetic inputs for the eskf_lio scan-to-map path.

Box-world scenes (ground + walls + pillars / buildings, or a tunnel), a spinning-LiDAR ray
caster with per-point sweep time (so that IMU deskew matters), and a 200 Hz IMU along a
planar arc.  Point records use the reference's wire format, pcl::PointXYZINormal (48 bytes):
x y z 1 | normal_x = time ratio, normal_y = ring, normal_z = sweep span [s], 0 | intensity,
curvature, pad, pad  (eskf_lio/src/feature_extract.cpp:337-346).

Measurement-side helper only: numpy, no GPU work, not part of the parity oracle.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

SEED_BASE = 0xDA1171
SWEEP = 0.1  # seconds per revolution
G = 9.81


@dataclass
class Scene:
    half: float                      # ground is [-half, half]^2 at z = 0
    boxes: np.ndarray                # (B, 6): xmin ymin zmin xmax ymax zmax
    ceiling: float | None = None     # z of a ceiling plane (tunnel) or None
    name: str = "box-world"


def make_box_world(half: float = 100.0, n_boxes: int = 48, seed: int = 1, keep_clear: float = 6.0, height: tuple = (2.0, 15.0),
                   max_size: float | None = None) -> Scene:
    """Ground + outer walls + n_boxes pillars/buildings; a corridor around the x axis stays free."""
    rng = np.random.default_rng(SEED_BASE + seed)
    boxes = []
    w = 1.0
    h = 12.0
    boxes.append([-half - w, -half - w, 0, -half, half + w, h])
    boxes.append([half, -half - w, 0, half + w, half + w, h])
    boxes.append([-half, -half - w, 0, half, -half, h])
    boxes.append([-half, half, 0, half, half + w, h])
    tries = 0
    while len(boxes) < 4 + n_boxes and tries < 100000:
        tries += 1
        cx, cy = rng.uniform(-half * 0.95, half * 0.95, 2)
        sx, sy = rng.uniform(1.0, (0.12 * half + 2.0) if max_size is None else max_size, 2)
        hh = rng.uniform(height[0], height[1])
        if abs(cy) - sy / 2 < keep_clear:
            continue
        boxes.append([cx - sx / 2, cy - sy / 2, 0, cx + sx / 2, cy + sy / 2, hh])
    return Scene(half=half, boxes=np.asarray(boxes, np.float64))


def make_tunnel(length: float = 400.0, width: float = 3.0, height: float = 3.0) -> Scene:
    """3 m x 3 m x 400 m tunnel along x (degenerate along x): floor z=0, ceiling, two side walls, end caps."""
    hl = length / 2
    hw = width / 2
    t = 1.0
    boxes = [
        [-hl, -hw - t, 0, hl, -hw, height],
        [-hl, hw, 0, hl, hw + t, height],
        [-hl - t, -hw - t, 0, -hl, hw + t, height],
        [hl, -hw - t, 0, hl + t, hw + t, height],
    ]
    return Scene(half=hl + t, boxes=np.asarray(boxes, np.float64), ceiling=height, name="tunnel")


def sample_map(scene: Scene, spacing: float = 0.5, jitter: float = 0.2, seed: int = 2, max_points: int | None = None,
               region: tuple | None = None) -> np.ndarray:
    """Surface samples at ~1 point per `spacing` voxel with U(-jitter, jitter) in-plane jitter. Returns (M,4) xyzi."""
    rng = np.random.default_rng(SEED_BASE + 100 + seed)
    pts = []

    def grid(u0, u1, v0, v1):
        us = np.arange(u0 + spacing / 2, u1, spacing)
        vs = np.arange(v0 + spacing / 2, v1, spacing)
        if len(us) == 0 or len(vs) == 0:
            return np.zeros((0, 2))
        uu, vv = np.meshgrid(us, vs, indexing="ij")
        g = np.stack([uu.ravel(), vv.ravel()], 1)
        g += rng.uniform(-jitter, jitter, g.shape)
        return g

    if scene.name == "tunnel":
        x0, x1 = scene.boxes[0][0], scene.boxes[0][3]
        y0, y1 = scene.boxes[0][4], scene.boxes[1][1]
        g = grid(x0, x1, y0, y1)
        pts.append(np.column_stack([g, rng.normal(0, 0.01, len(g))]))
        pts.append(np.column_stack([g + rng.uniform(-0.05, 0.05, g.shape), scene.ceiling + rng.normal(0, 0.01, len(g))]))
    else:
        x0 = y0 = -scene.half
        x1 = y1 = scene.half
        if region is not None:
            x0, x1, y0, y1 = region
        g = grid(x0, x1, y0, y1)
        inside = np.zeros(len(g), bool)
        for b in scene.boxes:
            inside |= (g[:, 0] > b[0]) & (g[:, 0] < b[3]) & (g[:, 1] > b[1]) & (g[:, 1] < b[4])
        g = g[~inside]
        pts.append(np.column_stack([g, rng.normal(0, 0.01, len(g))]))
    for b in scene.boxes:
        xmin, ymin, zmin, xmax, ymax, zmax = b
        for (fixed_axis, val, u0, u1) in ((0, xmin, ymin, ymax), (0, xmax, ymin, ymax), (1, ymin, xmin, xmax), (1, ymax, xmin, xmax)):
            g = grid(u0, u1, zmin, zmax)
            if len(g) == 0:
                continue
            nrm = val + rng.normal(0, 0.01, len(g))
            if fixed_axis == 0:
                p = np.column_stack([nrm, g[:, 0], g[:, 1]])
            else:
                p = np.column_stack([g[:, 0], nrm, g[:, 1]])
            pts.append(p)
        if scene.name != "tunnel" and zmax < 8:
            g = grid(xmin, xmax, ymin, ymax)
            if len(g):
                pts.append(np.column_stack([g, zmax + rng.normal(0, 0.01, len(g))]))
    P = np.concatenate(pts, 0)
    if scene.name == "tunnel":
        keep = (np.abs(P[:, 1]) <= scene.boxes[1][1] + 0.3) & (P[:, 2] >= -0.2) & (P[:, 2] <= scene.ceiling + 0.2)
        P = P[keep]
    if region is not None and scene.name != "tunnel":
        x0, x1, y0, y1 = region
        P = P[(P[:, 0] >= x0) & (P[:, 0] <= x1) & (P[:, 1] >= y0) & (P[:, 1] <= y1)]
    if max_points is not None and len(P) > max_points:
        idx = rng.permutation(len(P))[:max_points]
        P = P[np.sort(idx)]
    inten = rng.uniform(1.0, 100.0, len(P))
    return np.column_stack([P, inten]).astype(np.float32)


@dataclass
class Trajectory:
    """Planar arc at constant speed and yaw rate, sensor height z0."""
    speed: float = 1.0
    yaw_rate: float = 0.1
    x0: float = 0.0
    y0: float = 0.0
    z0: float = 1.5
    yaw0: float = 0.0

    def yaw(self, t):
        return self.yaw0 + self.yaw_rate * np.asarray(t, np.float64)

    def pos(self, t):
        t = np.asarray(t, np.float64)
        if abs(self.yaw_rate) < 1e-12:
            x = self.x0 + self.speed * np.cos(self.yaw0) * t
            y = self.y0 + self.speed * np.sin(self.yaw0) * t
        else:
            r = self.speed / self.yaw_rate
            x = self.x0 + r * (np.sin(self.yaw(t)) - np.sin(self.yaw0))
            y = self.y0 - r * (np.cos(self.yaw(t)) - np.cos(self.yaw0))
        return np.stack([x, y, np.full_like(x, self.z0)], -1)

    def vel(self, t):
        yw = self.yaw(t)
        return np.stack([self.speed * np.cos(yw), self.speed * np.sin(yw), np.zeros_like(yw)], -1)

    def acc(self, t):
        yw = self.yaw(t)
        a = self.speed * self.yaw_rate
        return np.stack([-a * np.sin(yw), a * np.cos(yw), np.zeros_like(yw)], -1)

    def rot(self, t):
        yw = self.yaw(t)
        c, s = np.cos(yw), np.sin(yw)
        R = np.zeros(np.shape(yw) + (3, 3))
        R[..., 0, 0] = c
        R[..., 0, 1] = -s
        R[..., 1, 0] = s
        R[..., 1, 1] = c
        R[..., 2, 2] = 1.0
        return R

    def pose24(self, t, R_L_I=None, T_L_I=None):
        R = self.rot(float(t))
        p = self.pos(float(t))
        RL = np.eye(3) if R_L_I is None else np.asarray(R_L_I, np.float64)
        TL = np.zeros(3) if T_L_I is None else np.asarray(T_L_I, np.float64)
        return np.concatenate([R.ravel(), p.ravel(), RL.ravel(), TL.ravel()])


def _raycast(scene: Scene, o: np.ndarray, d: np.ndarray, max_range: float) -> np.ndarray:
    """Nearest hit distance of rays (o + t d) against ground, ceiling and boxes; inf when nothing within range."""
    n = len(d)
    best = np.full(n, np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = -o[:, 2] / d[:, 2]
        ok = (d[:, 2] < 0) & (t > 0)
        hx = o[:, 0] + t * d[:, 0]
        hy = o[:, 1] + t * d[:, 1]
        ok &= (np.abs(hx) <= scene.half) & (np.abs(hy) <= scene.half)
        best = np.where(ok, t, best)
        if scene.ceiling is not None:
            t = (scene.ceiling - o[:, 2]) / d[:, 2]
            ok = (d[:, 2] > 0) & (t > 0)
            best = np.where(ok & (t < best), t, best)
        inv = 1.0 / d
        c = o.mean(0)
        reach = max_range + float(np.abs(o - c).max()) + 1.0
        for b in scene.boxes:
            if (b[0] - c[0] > reach or c[0] - b[3] > reach or b[1] - c[1] > reach or c[1] - b[4] > reach):
                continue
            t1 = (b[0:3] - o) * inv
            t2 = (b[3:6] - o) * inv
            tmin = np.minimum(t1, t2).max(1)
            tmax = np.maximum(t1, t2).min(1)
            hit = (tmax >= np.maximum(tmin, 0.0)) & (tmin > 1e-6)
            best = np.where(hit & (tmin < best), tmin, best)
    best[best > max_range] = np.inf
    return best


def simulate_scan(scene: Scene, traj: Trajectory, t_beg: float, beams: int, azimuths: int, fov_deg: tuple = (-15.0, 15.0),
                  seed: int = 3, max_range: float = 100.0, min_range: float = 0.5, range_sigma: float = 0.02,
                  sweep: float = SWEEP, motion: bool = True) -> np.ndarray:
    """One sweep of a spinning LiDAR starting at t_beg.  Returns (N, 12) float32 PointXYZINormal records in
    acquisition order (azimuth-major), points in the sensor frame AT THEIR OWN acquisition time."""
    rng = np.random.default_rng(SEED_BASE + 1000 + seed)
    el = np.deg2rad(np.linspace(fov_deg[0], fov_deg[1], beams))
    az_frac = np.arange(azimuths, dtype=np.float64) / azimuths
    az = 2 * np.pi * az_frac
    A, E = np.meshgrid(az, el, indexing="ij")            # azimuth-major
    ring = np.broadcast_to(np.arange(beams)[None, :], A.shape).ravel()
    frac = np.broadcast_to(az_frac[:, None], A.shape).ravel()
    A = A.ravel()
    E = E.ravel()
    d_body = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], 1)
    tt = t_beg + (frac * sweep if motion else np.zeros_like(frac))
    R = traj.rot(tt)
    o = traj.pos(tt)
    d_world = np.einsum("nij,nj->ni", R, d_body)
    rngs = _raycast(scene, o, d_world, max_range)
    ok = np.isfinite(rngs) & (rngs >= min_range)
    rngs = np.where(ok, rngs, 0.0) + rng.normal(0, range_sigma, len(rngs))
    p = d_body * rngs[:, None]
    out = np.zeros((int(ok.sum()), 12), np.float32)
    out[:, 0:3] = p[ok]
    out[:, 3] = 1.0
    out[:, 4] = frac[ok]                  # normal_x: time ratio in the sweep
    out[:, 5] = ring[ok]                  # normal_y: ring
    out[:, 6] = sweep                     # normal_z: sweep span [s]
    out[:, 8] = rng.uniform(1.0, 100.0, int(ok.sum()))  # intensity
    return out


def simulate_imu(traj: Trajectory, t0: float, t1: float, rate: float = 200.0, seed: int = 4, acc_sigma: float = 0.012,
                 gyr_sigma: float = 0.003) -> np.ndarray:
    """IMU samples with t0 < t <= t1 ... returns (m, 7): t, acc(3) [m/s^2 specific force, body], gyr(3)."""
    rng = np.random.default_rng(SEED_BASE + 2000 + seed)
    k0 = int(np.floor(t0 * rate)) + 1
    k1 = int(np.floor(t1 * rate + 1e-9))
    ts = np.arange(k0, k1 + 1, dtype=np.float64) / rate
    if len(ts) == 0:
        return np.zeros((0, 7))
    R = traj.rot(ts)
    a_w = traj.acc(ts) + np.array([0.0, 0.0, G])
    f_b = np.einsum("nji,nj->ni", R, a_w)
    gyr = np.tile(np.array([0.0, 0.0, traj.yaw_rate]), (len(ts), 1))
    f_b = f_b + rng.normal(0, acc_sigma, f_b.shape)
    gyr = gyr + rng.normal(0, gyr_sigma, gyr.shape)
    return np.column_stack([ts, f_b, gyr])


@dataclass
class ScanSpec:
    beams: int
    azimuths: int
    fov_deg: tuple
    max_range: float = 100.0


VLP16 = ScanSpec(16, 1800, (-15.0, 15.0))
OS1_64 = ScanSpec(64, 2048, (-22.5, 22.5))
OS1_64_1024 = ScanSpec(64, 1024, (-22.5, 22.5))
OS_32 = ScanSpec(32, 1024, (-22.5, 22.5))


@dataclass
class Sequence:
    """A replayable sequence: scene, trajectory, scan spec; yields (pts48, lidar_beg_time, imu7) per scan."""
    scene: Scene
    traj: Trajectory
    spec: ScanSpec
    seed: int = 0
    t_start: float = 0.0
    motion: bool = True
    cache: dict = field(default_factory=dict)

    def scan(self, k: int):
        if k in self.cache:
            return self.cache[k]
        t_beg = self.t_start + k * SWEEP
        pts = simulate_scan(self.scene, self.traj, t_beg, self.spec.beams, self.spec.azimuths, self.spec.fov_deg,
                            seed=self.seed * 1000 + k, max_range=self.spec.max_range, motion=self.motion)
        imu = simulate_imu(self.traj, t_beg - 1e-9 if k > 0 else t_beg - SWEEP, t_beg + SWEEP - 1e-9, seed=self.seed * 1000 + k)
        self.cache[k] = (pts, t_beg, imu)
        return self.cache[k]
