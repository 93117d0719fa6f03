"""Synthetic LiDAR frames with a planted intensity chessboard (host side, numpy).

Workload generator for BASELINE.json configs 2-5: a spinning multi-beam LiDAR is ray-cast
against a chessboard panel (squares of `square` m, white border), a vertical pillar next to it (clutter, so
that the ROI holds more than one cluster, as in the bundled bags), a floor and a cylindrical wall, so every
ray returns and a frame always has rings*azimuths points.  Range noise N(0, 0.01 m) along the ray;
intensities are integer-valued floats like a Velodyne driver emits (black N(20,s), white N(70,s), background
N(40,15), s~U(2,10), clamped to [0,255]).  The click is the true board centre + N(0, 0.05 m)^3.
Randomness is counter-based: Philox keyed by (base_seed, frame index), so any frame can be regenerated alone.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

BASE_SEED = 20181203


@dataclass(frozen=True)
class Sensor:
    name: str
    elevations_deg: tuple
    n_azimuth: int
    firing_order: tuple  # ring index emitted at each position of one azimuth step

    @property
    def n_rings(self):
        return len(self.elevations_deg)

    @property
    def n_points(self):
        return self.n_rings * self.n_azimuth


@dataclass(frozen=True)
class Board:
    corners_long: int   # inner corners along the long side (7)
    corners_short: int  # inner corners along the short side (5)
    square: float
    border: float = 0.10

    @property
    def cells(self):
        """(cells_x, cells_y) as ilcc2 sorts them: ascending."""
        a, b = self.corners_short + 1, self.corners_long + 1
        return (min(a, b), max(a, b))

    @property
    def n_corners(self):
        return self.corners_long * self.corners_short


def _vlp16_order():
    o = []
    for k in range(8):
        o += [k, k + 8]
    return tuple(o)


VLP16 = Sensor("vlp16", tuple(-15.0 + 2.0 * k for k in range(16)), 1800, _vlp16_order())
HDL64 = Sensor("hdl64", tuple(np.linspace(2.0, -24.8, 64).tolist()), 2048, tuple(range(64)))
BOARD_7x5 = Board(7, 5, 0.15)
BOARD_9x7 = Board(9, 7, 0.15)


@dataclass(frozen=True)
class Scene:
    sensor: Sensor
    board: Board
    range_min: float
    range_max: float
    floor_z: float
    wall_radius: float = 6.0
    center_z: tuple | None = (-0.1, 0.4)      # board-centre height range, or None -> use centre_elev_deg
    center_elev_deg: tuple | None = None
    pillar: bool = True


SCENE_VLP16 = Scene(VLP16, BOARD_7x5, 1.8, 3.2, -1.2)
SCENE_HDL64 = Scene(HDL64, BOARD_9x7, 2.0, 3.5, -2.0, center_z=None, center_elev_deg=(-18.0, -4.0))


def _rot(axis, ang):
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)


def _ray_dirs(sensor: Sensor):
    az = np.deg2rad(-180.0 + (360.0 / sensor.n_azimuth) * np.arange(sensor.n_azimuth))
    el = np.deg2rad(np.asarray(sensor.elevations_deg))[list(sensor.firing_order)]
    ring = np.asarray(sensor.firing_order, np.uint16)
    azg, elg = np.meshgrid(az, el, indexing="ij")  # azimuth-major, rings interleaved in firing order
    d = np.stack([np.cos(elg) * np.cos(azg), np.cos(elg) * np.sin(azg), np.sin(elg)], -1).reshape(-1, 3)
    rings = np.broadcast_to(ring[None, :], azg.shape).reshape(-1)
    return d, rings


_DIR_CACHE = {}


def make_frame(index: int, scene: Scene = SCENE_VLP16, base_seed: int = BASE_SEED):
    """Returns (points (N,4) float32 XYZI, click (3,) float32, truth dict)."""
    rng = np.random.Generator(np.random.Philox(key=[base_seed, index]))
    s, b = scene.sensor, scene.board
    if s.name not in _DIR_CACHE:
        _DIR_CACHE[s.name] = _ray_dirs(s)
    d, rings = _DIR_CACHE[s.name]
    n = d.shape[0]

    R = rng.uniform(scene.range_min, scene.range_max)
    az = np.deg2rad(rng.uniform(-30.0, 30.0))
    if scene.center_z is not None:
        cz = rng.uniform(*scene.center_z)
    else:
        cz = R * np.tan(np.deg2rad(rng.uniform(*scene.center_elev_deg)))
    c = np.array([R * np.cos(az), R * np.sin(az), cz])
    yaw, pitch = np.deg2rad(rng.uniform(-25.0, 25.0, 2))
    roll = np.deg2rad(rng.uniform(-15.0, 15.0))
    sigma_i = rng.uniform(2.0, 10.0)
    phase = int(rng.integers(0, 2))
    side = 1.0 if rng.integers(0, 2) else -1.0
    gap = rng.uniform(0.2, 0.5)

    n0 = -np.array([np.cos(az), np.sin(az), 0.0])            # facing the sensor
    u0 = np.cross(np.array([0.0, 0.0, 1.0]), n0)             # horizontal, in plane (long side)
    v0 = np.cross(n0, u0)                                    # up (short side)
    Rm = _rot(n0, roll) @ _rot(u0, pitch) @ _rot(np.array([0.0, 0.0, 1.0]), yaw)
    nrm, u, v = Rm @ n0, Rm @ u0, Rm @ v0
    half_u = b.square * (b.corners_long + 1) / 2.0
    half_v = b.square * (b.corners_short + 1) / 2.0

    # background: wall and floor
    ch = np.hypot(d[:, 0], d[:, 1])
    t = scene.wall_radius / ch
    with np.errstate(divide="ignore"):
        tf = np.where(d[:, 2] < 0, scene.floor_z / d[:, 2], np.inf)
    t = np.minimum(t, tf)
    kind = np.zeros(n, np.int8)  # 0 background, 1 black, 2 white
    # pillar
    if scene.pillar:
        pr = 0.12
        pc = c + u * side * (half_u + b.border + pr + gap)
        a2 = ch * ch
        bq = d[:, 0] * pc[0] + d[:, 1] * pc[1]
        cq = pc[0] ** 2 + pc[1] ** 2 - pr * pr
        disc = bq * bq - a2 * cq
        tp = np.where(disc > 0, (bq - np.sqrt(np.maximum(disc, 0))) / a2, np.inf)
        zp = tp * d[:, 2]
        hit = (tp > 0) & (tp < t) & (zp <= cz + 0.9)
        t = np.where(hit, tp, t)
    # board
    dn = d @ nrm
    with np.errstate(divide="ignore", invalid="ignore"):
        tb = (c @ nrm) / dn
    p = tb[:, None] * d - c
    a, bb = p @ u, p @ v
    on = (tb > 0) & (tb < t) & (np.abs(a) <= half_u + b.border) & (np.abs(bb) <= half_v + b.border)
    t = np.where(on, tb, t)
    inner = on & (np.abs(a) < half_u) & (np.abs(bb) < half_v)
    ia = np.floor((a + half_u) / b.square).astype(np.int64)
    ib = np.floor((bb + half_v) / b.square).astype(np.int64)
    black = inner & (((ia + ib + phase) & 1) == 0)
    kind[on] = 2
    kind[black] = 1

    t = t + rng.normal(0.0, 0.01, n)
    xyz = t[:, None] * d
    mean = np.choose(kind, [40.0, 20.0, 70.0])
    sd = np.choose(kind, [15.0, sigma_i, sigma_i])
    inten = np.rint(np.clip(mean + sd * rng.standard_normal(n), 0.0, 255.0))
    pts = np.empty((n, 4), np.float32)
    pts[:, :3] = xyz
    pts[:, 3] = inten
    click = (c + rng.normal(0.0, 0.05, 3)).astype(np.float32)

    gl = []
    for i in range(1, b.corners_long + 1):
        for j in range(1, b.corners_short + 1):
            gl.append(c + u * (i * b.square - half_u) + v * (j * b.square - half_v))
    truth = dict(center=c, normal=nrm, u=u, v=v, corners=np.asarray(gl), n_board=int(on.sum()), sigma_i=sigma_i,
                 phase=phase, rings=rings)
    return pts, click, truth


def make_batch(n_frames: int, start: int = 0, scene: Scene = SCENE_VLP16, base_seed: int = BASE_SEED, out=None,
               clicks_out=None, want_truth: bool = False):
    """(points (F,N,4) float32, clicks (F,3) float32[, truth corners (F,ncorn,3)]) for frames start..start+F-1."""
    N = scene.sensor.n_points
    pts = out if out is not None else np.empty((n_frames, N, 4), np.float32)
    clicks = clicks_out if clicks_out is not None else np.empty((n_frames, 3), np.float32)
    gt = np.empty((n_frames, scene.board.n_corners, 3)) if want_truth else None
    for f in range(n_frames):
        p, c, tr = make_frame(start + f, scene, base_seed)
        pts[f] = p
        clicks[f] = c
        if want_truth:
            gt[f] = tr["corners"]
    if want_truth:
        return pts, clicks, gt
    return pts, clicks


def corner_set_distance(a: np.ndarray, b: np.ndarray) -> float:
    """Order-free parity metric: max over corners of a of the distance to the nearest corner of b (and vice versa)."""
    d = np.linalg.norm(a[:, None, :].astype(np.float64) - b[None, :, :].astype(np.float64), axis=2)
    return float(max(d.min(1).max(), d.min(0).max()))


# ---- synthetic camera side for BASELINE.json configs[4] (SURVEY.md 8(d) "Config 5") ----
# Planted LiDAR->camera extrinsic = the reference's shipped ilcc2/config/pointgrey.bin (row-major 4x4) and the
# intrinsics of ilcc2/config/pointgrey.yaml (fx, cx, fy, cy).
PLANTED_EXTRINSIC = np.array([
    [0.030776439680377833, -0.999522902318748, 0.0026035554356903844, -0.05487600789921265],
    [-0.003282631634222143, -0.0027058505464300914, -0.9999909513102478, -0.0740006228856876],
    [0.999520902778003, 0.030767614680488886, -0.003364341944407353, -0.08264568089777906],
    [0.0, 0.0, 0.0, 1.0]])
PLANTED_CAMERA = np.array([1061.37439737547, 980.706836288949, 1061.02435228316, 601.685030610243])


def project_truth(gt: np.ndarray, scene: Scene = SCENE_VLP16, start: int = 0, noise_px: float = 0.5,
                  base_seed: int = BASE_SEED) -> np.ndarray:
    """Image corners (F, ncorn, 2) of the ground-truth 3-D corners under the planted extrinsic + N(0, noise_px),
    laid out like the reference's camera corner files after read_cam_corners: rows along the short board side,
    columns along the long one (the LiDAR lattice of getPCDcorners has the same shape; the reference's
    check_order_lidar / check_order_cam remove the remaining flips)."""
    b = scene.board
    F = gt.shape[0]
    q = gt.reshape(F, b.corners_long, b.corners_short, 3).transpose(0, 2, 1, 3).reshape(F, -1, 3)
    q = q @ PLANTED_EXTRINSIC[:3, :3].T + PLANTED_EXTRINSIC[:3, 3]
    fx, cx, fy, cy = PLANTED_CAMERA
    uv = np.stack([fx * q[..., 0] / q[..., 2] + cx, fy * q[..., 1] / q[..., 2] + cy], -1)
    for f in range(F):
        rng = np.random.Generator(np.random.Philox(key=[base_seed ^ 0x5EED, start + f]))
        uv[f] += rng.normal(0.0, noise_px, uv[f].shape)
    return uv
