"""Synthetic inputs of the shapes BASELINE.json names.

Host generators (numpy) restate the reference demos' input recipes without
pygame so they can be seeded and replayed headlessly:

* :func:`config1_pair`   - visualization.py:9-44,169-194 (circle + square pair)
* :class:`LidarSim`      - robot-visualization.py:11-40,42-120,222-237 (ray caster)
* :func:`patches3d_pair` - 3-D planar-patch clouds (configs 3-5; ours, the
  reference has no 3-D inputs)

:func:`patches3d_batch_device` builds the config-4 batch directly in HBM with
torch (device plumbing only - no registration math lives here).
"""
from __future__ import annotations

import math
import random

import numpy as np


# --------------------------------------------------------------------------
# config 1: visualization.py
# --------------------------------------------------------------------------
def _square_points(center, size, per_side, rnd):
    """visualization.py:9-30 (random.sample keeps half of the 4*per_side points)."""
    half = size / 2
    xs = np.linspace(center[0] - half, center[0] + half, per_side)
    ys = np.linspace(center[1] - half, center[1] + half, per_side)
    pts = [[x, center[1] - half] for x in xs]
    pts += [[center[0] + half, y] for y in ys]
    pts += [[x, center[1] + half] for x in xs]
    pts += [[center[0] - half, y] for y in ys]
    keep = rnd.sample(range(len(pts)), len(pts) // 2)
    return np.array([pts[i] for i in range(len(pts)) if i in keep])


def _circle_points(center, radius, count, rnd):
    """visualization.py:32-39."""
    pts = []
    for _ in range(count):
        a = rnd.uniform(0, 2 * np.pi)
        pts.append([center[0] + radius * np.cos(a), center[1] + radius * np.sin(a)])
    return np.array(pts)


def config1_pair(seed):
    """visualization.py:169-194 with random.seed(seed); np.random.seed(seed).
    The circle is drawn before the square, exactly in the script's call order,
    so a seeded run of the original script yields the same arrays.
    Returns (source (90,2), target (87,2)) float64."""
    rnd = random.Random(seed)
    nrs = np.random.RandomState(seed)
    circle = _circle_points((300, 150), 100, 30, rnd)
    square = _square_points((600, 250), 200, 30, rnd)
    src = np.concatenate([circle, square])
    ang = np.pi / 3
    rot = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
    tgt = np.dot(src, rot.T) + np.array([150, -50])        # visualization.py:42-44
    src = src + nrs.normal(0, 2, src.shape)
    tgt = tgt + nrs.normal(0, 5, tgt.shape)
    nrs.shuffle(tgt)
    tgt = tgt[:len(tgt) - 3]
    return src, tgt


# --------------------------------------------------------------------------
# config 2: robot-visualization.py ray caster
# --------------------------------------------------------------------------
class Rect:
    """Minimal stand-in for pygame.Rect (robot-visualization.py:36-37,52-57)."""

    def __init__(self, x, y, w, h):
        self.topleft = (x, y)
        self.topright = (x + w, y)
        self.bottomleft = (x, y + h)
        self.bottomright = (x + w, y + h)


def _ray_line(p1, p2, p3, p4):
    """robot-visualization.py:79-93."""
    x1, y1 = p1
    x2, y2 = p2
    x3, y3 = p3
    x4, y4 = p4
    denom = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4)
    if denom == 0:
        return None
    t = ((x1 - x3) * (y3 - y4) - (y1 - y3) * (x3 - x4)) / denom
    u = -((x1 - x2) * (y1 - y3) - (y1 - y2) * (x1 - x3)) / denom
    if 0 <= t <= 1 and 0 <= u <= 1:
        return (x1 + t * (x2 - x1), y1 + t * (y2 - y1))
    return None


def _ray_circle(p1, p2, center, radius):
    """robot-visualization.py:95-120."""
    x1, y1 = p1
    x2, y2 = p2
    cx, cy = center
    dx, dy = x2 - x1, y2 - y1
    fx, fy = x1 - cx, y1 - cy
    a = dx * dx + dy * dy
    b = 2 * (fx * dx + fy * dy)
    c = (fx * fx + fy * fy) - radius * radius
    disc = b * b - 4 * a * c
    if disc >= 0:
        disc = math.sqrt(disc)
        t1 = (-b - disc) / (2 * a)
        t2 = (-b + disc) / (2 * a)
        hits = []
        if 0 <= t1 <= 1:
            hits.append((x1 + t1 * dx, y1 + t1 * dy))
        if 0 <= t2 <= 1:
            hits.append((x1 + t2 * dx, y1 + t2 * dy))
        return hits if hits else None
    return None


class LidarSim:
    """Headless replay of the robot demo's world, ray caster and key handling.

    Constants from robot-visualization.py:19-26,35-40.  ``num_rays`` is 90 in
    the reference (line 22); 360 reproduces BASELINE.json's wording."""

    MAX_RAY_RANGE = 400
    ROBOT_SPEED = 2
    ROBOT_YAW_SPEED = 2
    NOISE = 2

    def __init__(self, seed=0, num_rays=90):
        self.rnd = random.Random(seed)
        self.num_rays = num_rays
        self.x, self.y, self.yaw = 50.0, 400.0, 0
        self.obstacles = [Rect(100, 250, 200, 50), Rect(400, 450, 50, 200), (600, 300, 50), (200, 550, 75)]

    def cast_ray(self, angle):
        """robot-visualization.py:42-77."""
        x1, y1 = self.x, self.y
        x2 = x1 + self.MAX_RAY_RANGE * math.cos(math.radians(angle))
        y2 = y1 + self.MAX_RAY_RANGE * math.sin(math.radians(angle))
        best, hit = float("inf"), None
        for ob in self.obstacles:
            if isinstance(ob, Rect):
                for a, b in [(ob.topleft, ob.topright), (ob.topright, ob.bottomright),
                             (ob.bottomright, ob.bottomleft), (ob.bottomleft, ob.topleft)]:
                    p = _ray_line((x1, y1), (x2, y2), a, b)
                    if p:
                        d = math.hypot(p[0] - x1, p[1] - y1)
                        if d < best:
                            best, hit = d, p
            else:
                pts = _ray_circle((x1, y1), (x2, y2), (ob[0], ob[1]), ob[2])
                if pts:
                    for p in pts:
                        d = math.hypot(p[0] - x1, p[1] - y1)
                        if d < best:
                            best, hit = d, p
        if hit:
            return best + self.rnd.uniform(-self.NOISE, self.NOISE)
        return None

    def step(self, left=False, right=False, up=False, down=False):
        """One UI tick of key handling, robot-visualization.py:210-220."""
        if left:
            self.yaw -= self.ROBOT_YAW_SPEED
        if right:
            self.yaw += self.ROBOT_YAW_SPEED
        if up:
            self.x += self.ROBOT_SPEED * math.cos(math.radians(self.yaw))
            self.y += self.ROBOT_SPEED * math.sin(math.radians(self.yaw))
        if down:
            self.x -= self.ROBOT_SPEED * math.cos(math.radians(self.yaw))
            self.y -= self.ROBOT_SPEED * math.sin(math.radians(self.yaw))

    def scan(self):
        """robot-visualization.py:222-237: list of robot-relative hit points."""
        pts = []
        for angle in range(self.yaw, self.yaw + 360, 360 // self.num_rays):
            d = self.cast_ray(angle)
            if d:
                pts.append((d * math.cos(math.radians(angle - self.yaw)),
                            d * math.sin(math.radians(angle - self.yaw))))
        return pts


def lidar_sequence(seed=0, num_rays=90, n_scans=30, ticks_per_scan=5):
    """Scripted drive: blocks of 6 scans straight / turning right / turning
    left while moving forward.  Returns (scans, poses):
    scans[i] is a list of (x, y) tuples exactly as gicp_worker receives them
    (robot-visualization.py:155-156), poses[i] = (x, y, yaw_deg)."""
    sim = LidarSim(seed, num_rays)
    scans, poses = [], []
    for s in range(n_scans):
        block = (s // 6) % 3
        for _ in range(ticks_per_scan):
            sim.step(up=True, right=(block == 1), left=(block == 2))
        scans.append(sim.scan())
        poses.append((sim.x, sim.y, sim.yaw))
    return scans, poses


# --------------------------------------------------------------------------
# configs 3-5: 3-D planar patches
# --------------------------------------------------------------------------
def _rand_rotation(rng, max_deg):
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    ang = np.deg2rad(rng.uniform(-max_deg, max_deg))
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


def patches3d_pair(n=100_000, n_patches=64, cube=100.0, patch=30.0, sigma=0.02,
                   max_rot_deg=2.0, max_trans=0.5, seed=0):
    """Source and target drawn independently from the same random planar
    patches; the target surface is moved by a rigid motion about the cube
    centre.  Returned as float32 (N,3) arrays plus the 4x4 ground truth that
    maps source onto target."""
    rng = np.random.default_rng(seed)
    centres = rng.uniform(0.25 * cube, 0.75 * cube, size=(n_patches, 3))
    frames = np.stack([np.linalg.qr(rng.normal(size=(3, 3)))[0] for _ in range(n_patches)])

    def sample():
        which = rng.integers(0, n_patches, size=n)
        ab = rng.uniform(-patch / 2, patch / 2, size=(n, 2))
        p = centres[which] + ab[:, :1] * frames[which, :, 0] + ab[:, 1:] * frames[which, :, 1]
        return p + rng.normal(0, sigma, size=(n, 3))

    src = sample()
    tgt = sample()
    R = _rand_rotation(rng, max_rot_deg)
    tdir = rng.normal(size=3)
    t = tdir / np.linalg.norm(tdir) * rng.uniform(0, max_trans)
    c = np.full(3, cube / 2)
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = c - R @ c + t
    tgt = tgt @ R.T + T[:3, 3]
    return src.astype(np.float32), tgt.astype(np.float32), T


CONFIG3 = dict(n=100_000, n_patches=64, cube=100.0, patch=30.0, sigma=0.02, max_rot_deg=2.0, max_trans=0.5)
CONFIG3_PARAMS = dict(k=20, max_distance_nearest_neighbors=6.0, max_distance_correspondence=4.0, tolerance=1e-6)
CONFIG4 = dict(n=32_768, n_patches=16, cube=40.0, patch=30.0, sigma=0.02, max_rot_deg=2.0, max_trans=0.5)
CONFIG4_PARAMS = dict(k=20, max_distance_nearest_neighbors=5.0, max_distance_correspondence=2.0, tolerance=1e-6)
CONFIG5 = dict(n=16_777_216, n_patches=1024, cube=400.0, patch=30.0, sigma=0.02, max_rot_deg=0.25, max_trans=0.5)
CONFIG5_PARAMS = dict(k=20, max_distance_nearest_neighbors=1.8, max_distance_correspondence=2.5, tolerance=1e-6)


def patches3d_batch_device(n_pairs, n=32_768, n_patches=16, cube=40.0, patch=30.0, sigma=0.02,
                           max_rot_deg=2.0, max_trans=0.5, seed=0, device="cuda", chunk=256):
    """Config-4 batch generated in HBM: (src, tgt) float32 (n_pairs*n, 3), CSR
    offsets (n_pairs+1,) int64 (identical for both sides) and the ground-truth
    4x4 per pair.  Same distribution as :func:`patches3d_pair`; the random
    stream is torch's Philox generator seeded with ``seed`` (per-chunk), so a
    sample of pairs can be copied back for the CPU oracle."""
    import torch

    g = torch.Generator(device=device)
    src = torch.empty((n_pairs * n, 3), dtype=torch.float32, device=device)
    tgt = torch.empty_like(src)
    Ts = torch.empty((n_pairs, 4, 4), dtype=torch.float64, device=device)
    for c0 in range(0, n_pairs, chunk):
        b = min(chunk, n_pairs - c0)
        g.manual_seed(seed * 1_000_003 + c0)
        centres = torch.empty((b, n_patches, 3), device=device).uniform_(0.25 * cube, 0.75 * cube, generator=g)
        frames, _ = torch.linalg.qr(torch.randn((b, n_patches, 3, 3), device=device, generator=g))
        axis = torch.randn((b, 3), device=device, dtype=torch.float64, generator=g)
        axis = axis / axis.norm(dim=1, keepdim=True)
        ang = torch.deg2rad(torch.empty((b,), device=device, dtype=torch.float64).uniform_(-max_rot_deg, max_rot_deg, generator=g))
        K = torch.zeros((b, 3, 3), device=device, dtype=torch.float64)
        K[:, 0, 1], K[:, 0, 2], K[:, 1, 0] = -axis[:, 2], axis[:, 1], axis[:, 2]
        K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -axis[:, 0], -axis[:, 1], axis[:, 0]
        eye = torch.eye(3, device=device, dtype=torch.float64).expand(b, 3, 3)
        R = eye + torch.sin(ang)[:, None, None] * K + (1 - torch.cos(ang))[:, None, None] * (K @ K)
        tdir = torch.randn((b, 3), device=device, dtype=torch.float64, generator=g)
        t = tdir / tdir.norm(dim=1, keepdim=True) * torch.empty((b, 1), device=device, dtype=torch.float64).uniform_(0, max_trans, generator=g)
        cc = torch.full((3,), cube / 2, device=device, dtype=torch.float64)
        tt = cc - (R @ cc) + t
        Ts[c0:c0 + b] = torch.eye(4, device=device, dtype=torch.float64)
        Ts[c0:c0 + b, :3, :3] = R
        Ts[c0:c0 + b, :3, 3] = tt

        def sample():
            which = torch.randint(0, n_patches, (b, n), device=device, generator=g)
            ab = torch.empty((b, n, 2), device=device).uniform_(-patch / 2, patch / 2, generator=g)
            ctr = torch.gather(centres, 1, which[..., None].expand(b, n, 3))
            fr = torch.gather(frames.reshape(b, n_patches, 9), 1, which[..., None].expand(b, n, 9)).reshape(b, n, 3, 3)
            p = ctr + ab[..., :1] * fr[..., :, 0] + ab[..., 1:] * fr[..., :, 1]
            return p + sigma * torch.randn((b, n, 3), device=device, generator=g)

        s = sample()
        q = sample().double()
        q = q @ R.transpose(1, 2) + tt[:, None, :]
        src[c0 * n:(c0 + b) * n] = s.reshape(-1, 3)
        tgt[c0 * n:(c0 + b) * n] = q.float().reshape(-1, 3)
    off = torch.arange(0, n_pairs + 1, dtype=torch.int64, device=device) * n
    return src, tgt, off, Ts
