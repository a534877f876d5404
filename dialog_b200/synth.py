"""Deterministic synthetic multi-plane scenes for the BASELINE.json configs.

Counter-based generation: every random number of point i is splitmix64(seed, i, stream), so any
index range can be generated independently (ranks generate their own shard) and the bytes do not
depend on chunking.  Classes (plane k / outlier) are assigned by hash, so points of different planes
are interleaved in index order, like a real scan after registration — not grouped.

Units are metres; sigma = 0.03 along the plane normal; the reference's distance threshold is 0.1
(Dialog/config.txt:29, T_dist_point_plane) and its minimum plane size 500 (Dialog/config.txt:20).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _u01(seed: int, idx: np.ndarray, stream: int) -> np.ndarray:
    """Uniform doubles in (0,1) for (seed, index, stream)."""
    with np.errstate(over="ignore"):
        key = _splitmix64(np.uint64(seed) ^ (np.uint64(stream) * np.uint64(0xD1B54A32D192ED03)))
        h = _splitmix64(idx.astype(np.uint64) * np.uint64(0x2545F4914F6CDD1D) ^ key)
    return ((h >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


@dataclass
class Patch:
    """Rectangular planar patch: origin + s*u + t*v, s,t in [0,1]."""
    origin: tuple
    u: tuple
    v: tuple
    share: float

    @property
    def normal(self) -> np.ndarray:
        n = np.cross(np.asarray(self.u, float), np.asarray(self.v, float))
        return n / np.linalg.norm(n)

    @property
    def coeff(self) -> np.ndarray:
        n = self.normal
        return np.array([n[0], n[1], n[2], -float(n @ np.asarray(self.origin, float))])


    def border(self, per_side: int = 45) -> np.ndarray:
        """The patch outline as a closed polygon of 4 * per_side vertices ((n,4) float32, w = 1) — what
        pcl::ConcaveHull would hand back as Plane::border for this patch (the reference's saved polygon sets hold
        ~180 vertices per polygon, Dialog/dataForPlane)."""
        o, u, v = (np.asarray(a, float) for a in (self.origin, self.u, self.v))
        t = np.arange(per_side) / per_side
        pts = np.concatenate([o + t[:, None] * u, o + u + t[:, None] * v, o + u + v - t[:, None] * u, o + v - t[:, None] * v])
        out = np.ones((pts.shape[0], 4), np.float32)
        out[:, :3] = pts.astype(np.float32)
        return out


@dataclass
class Scene:
    patches: list
    box_lo: tuple
    box_hi: tuple
    outlier_share: float
    sigma: float
    seed: int

    def points(self, start: int, stop: int, chunk: int = 1 << 22) -> np.ndarray:
        """Points [start, stop) as (n,4) float32 rows (x, y, z, 1) == pcl::PointXYZ."""
        out = np.empty((stop - start, 4), np.float32)
        for a in range(start, stop, chunk):
            b = min(stop, a + chunk)
            out[a - start: b - start] = self._chunk(np.arange(a, b, dtype=np.uint64))
        return out

    def labels(self, start: int, stop: int) -> np.ndarray:
        """Generating class per point: patch index, or -1 for outliers."""
        idx = np.arange(start, stop, dtype=np.uint64)
        return self._classes(idx)

    def _classes(self, idx: np.ndarray) -> np.ndarray:
        shares = np.array([p.share for p in self.patches], float)
        edges = np.cumsum(shares)
        assert edges[-1] + self.outlier_share <= 1.0 + 1e-9
        c = _u01(self.seed, idx, 0)
        k = np.searchsorted(edges, c, side="right").astype(np.int32)
        k[k >= len(self.patches)] = -1
        return k

    def _chunk(self, idx: np.ndarray) -> np.ndarray:
        n = idx.size
        k = self._classes(idx)
        s = _u01(self.seed, idx, 1)
        t = _u01(self.seed, idx, 2)
        g1 = _u01(self.seed, idx, 3)
        g2 = _u01(self.seed, idx, 4)
        gauss = np.sqrt(-2.0 * np.log(g1)) * np.cos(2.0 * np.pi * g2)  # Box-Muller
        lo = np.asarray(self.box_lo, float)
        hi = np.asarray(self.box_hi, float)
        # outliers: uniform in the box (streams 1, 2, 5)
        w = _u01(self.seed, idx, 5)
        xyz = lo + np.stack([s, t, w], axis=1) * (hi - lo)
        for j, p in enumerate(self.patches):
            m = k == j
            if not m.any():
                continue
            o = np.asarray(p.origin, float)
            u = np.asarray(p.u, float)
            v = np.asarray(p.v, float)
            xyz[m] = o + s[m, None] * u + t[m, None] * v + (self.sigma * gauss[m])[:, None] * p.normal
        out = np.ones((n, 4), np.float32)
        out[:, :3] = xyz.astype(np.float32)
        return out


def _unit_from_hash(seed: int, j: int) -> np.ndarray:
    a = _u01(seed, np.array([j], np.uint64), 90)[0]
    b = _u01(seed, np.array([j], np.uint64), 91)[0]
    z = 2.0 * a - 1.0
    phi = 2.0 * np.pi * b
    r = np.sqrt(max(0.0, 1.0 - z * z))
    return np.array([r * np.cos(phi), r * np.sin(phi), z])


def three_planes_scene(seed: int = 20260002, sigma: float = 0.03) -> Scene:
    """BASELINE config 2: box [0,3]^3, 3 patches with seeded random normals (shares 30/25/15 %),
    N(0, sigma^2) along the normal, 30 % uniform outliers."""
    patches = []
    shares = (0.30, 0.25, 0.15)
    centres = ((1.5, 1.5, 0.8), (1.2, 1.6, 1.6), (1.7, 1.3, 2.3))
    for j, (sh, c) in enumerate(zip(shares, centres)):
        n = _unit_from_hash(seed, j)
        a = np.cross(n, [0.0, 0.0, 1.0] if abs(n[2]) < 0.9 else [1.0, 0.0, 0.0])
        a /= np.linalg.norm(a)
        b = np.cross(n, a)
        ext = 2.0
        o = np.asarray(c) - 0.5 * ext * a - 0.5 * ext * b
        patches.append(Patch(tuple(o), tuple(ext * a), tuple(ext * b), sh))
    return Scene(patches, (0.0, 0.0, 0.0), (3.0, 3.0, 3.0), 0.30, sigma, seed)


def indoor_scene(seed: int = 20260003, sigma: float = 0.03) -> Scene:
    """BASELINE configs 3/4: a 30 x 20 x 3 m storey — floor, ceiling, 4 outer walls, 14 inner wall
    patches (axis-aligned and oblique), plane shares descending geometrically (sum 80 %), 20 % clutter."""
    L, W, H = 30.0, 20.0, 3.0
    geo = [
        ((0, 0, 0), (L, 0, 0), (0, W, 0)),            # floor
        ((0, 0, H), (L, 0, 0), (0, W, 0)),            # ceiling
        ((0, 0, 0), (L, 0, 0), (0, 0, H)),            # wall y = 0
        ((0, W, 0), (L, 0, 0), (0, 0, H)),            # wall y = W
        ((0, 0, 0), (0, W, 0), (0, 0, H)),            # wall x = 0
        ((L, 0, 0), (0, W, 0), (0, 0, H)),            # wall x = L
        ((6, 0, 0), (0, 12, 0), (0, 0, H)),           # inner, x = 6
        ((12, 4, 0), (0, 16, 0), (0, 0, H)),          # inner, x = 12
        ((18, 0, 0), (0, 11, 0), (0, 0, H)),          # inner, x = 18
        ((24, 6, 0), (0, 14, 0), (0, 0, H)),          # inner, x = 24
        ((0, 7, 0), (5, 0, 0), (0, 0, H)),            # inner, y = 7
        ((7, 13, 0), (10, 0, 0), (0, 0, H)),          # inner, y = 13
        ((19, 5, 0), (10, 0, 0), (0, 0, H)),          # inner, y = 5
        ((13, 16, 0), (10, 0, 0), (0, 0, H)),         # inner, y = 16
        ((1, 14, 0), (4, 5, 0), (0, 0, H)),           # oblique
        ((7, 1, 0), (4, 3, 0), (0, 0, H)),            # oblique
        ((13, 1, 0), (4, -0.8, 0), (0, 0, H)),        # oblique (shallow)
        ((25, 1, 0), (4, 4, 0), (0, 0, H)),           # oblique 45 deg
        ((20, 12, 0.2), (3, 0, 1.2), (0, 3, 0)),      # tilted slab (ramp)
        ((2, 2, 0.6), (2.5, 0, 0.5), (0, 2.5, 0.3)),  # tilted table top
    ]
    r = (1.5 / 12.0) ** (1.0 / 19.0)
    raw = np.array([r ** j for j in range(len(geo))])
    shares = 0.80 * raw / raw.sum()
    patches = [Patch(o, u, v, float(s)) for (o, u, v), s in zip(geo, shares)]
    return Scene(patches, (0.0, 0.0, 0.0), (L, W, H), 0.20, sigma, seed)


def tile_scene(cloud_id: int, seed: int = 20260005, sigma: float = 0.03) -> Scene:
    """BASELINE config 5: one small per-scan tile — box [0,3]^3, 2 planes (50/30 %) + 20 % outliers."""
    s = (seed ^ (cloud_id * 0x9E3779B1)) & 0x7FFFFFFFFFFFFFFF
    patches = []
    for j, (sh, c) in enumerate(((0.50, (1.5, 1.5, 1.0)), (0.30, (1.5, 1.5, 2.0)))):
        n = _unit_from_hash(s, j)
        a = np.cross(n, [0.0, 0.0, 1.0] if abs(n[2]) < 0.9 else [1.0, 0.0, 0.0])
        a /= np.linalg.norm(a)
        b = np.cross(n, a)
        o = np.asarray(c) - a - b
        patches.append(Patch(tuple(o), tuple(2.0 * a), tuple(2.0 * b), sh))
    return Scene(patches, (0.0, 0.0, 0.0), (3.0, 3.0, 3.0), 0.20, sigma, s)
