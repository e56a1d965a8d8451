"""Seeded, headless restatements of the reference demos' INPUT recipes.  TEST INFRASTRUCTURE ONLY
(it lives beside the tests, not in the product package): the demos cannot be imported without pygame,
and the fixtures under tests/golden/ must be regenerated bit for bit, so the recipes are restated here.

* :func:`config1_pair`   - visualization.py:9-44,169-194 (circle + square pair)
* :class:`LidarSim`      - robot-visualization.py:11-40,42-120,222-237 (ray caster, key handling)
* :func:`lidar_sequence` - a scripted drive through the demo's world
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
