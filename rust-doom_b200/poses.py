"""Synthetic camera-pose generators for the BASELINE.json configs.

* flythrough_poses : C2/C3/C4 -- Catmull-Rom spline through the centroids of a seeded random walk over
                     subsectors that are adjacent through two-sided segs; eye = floor + 41, yaw along
                     the tangent, pitch 0.  Evaluated in double on the host, then quantised to 16.16 / BAM.
* random_poses     : C5 -- splitmix64; position uniform in the level's bounding box, accepted iff inside a
                     subsector per the reference's `sector_at` rule (wad/src/visitor.rs:1028-1060) and the
                     sector is at least 56 high; yaw uniform over 2^32 BAM; eye = floor + 41.
"""
from __future__ import annotations

import math
import struct
from typing import List

import numpy as np

from . import POSE_DTYPE, Scene
from .synthwad import SplitMix64

EYE_HEIGHT = 41
MIN_ROOM_HEIGHT = 56


def _blob_arrays(blob: bytes):
    h = struct.unpack_from("<32I", blob, 0)
    verts = np.frombuffer(blob, dtype="<i4", count=h[3] * 2, offset=h[10]).reshape(-1, 2)
    ssec = np.frombuffer(blob, dtype="<i4", count=h[5] * 4, offset=h[12]).reshape(-1, 4)
    segs = np.frombuffer(blob, dtype="<i4", count=h[6] * 16, offset=h[13]).reshape(-1, 16)
    return verts, ssec, segs


def random_poses(scene: Scene, n: int, seed: int = 5) -> np.ndarray:
    verts, _, _ = _blob_arrays(scene.blob)
    x0, x1 = int(verts[:, 0].min()), int(verts[:, 0].max())
    y0, y1 = int(verts[:, 1].min()), int(verts[:, 1].max())
    rng = SplitMix64(seed)
    out = np.zeros(n, dtype=POSE_DTYPE)
    k = 0
    guard = 0
    while k < n:
        guard += 1
        if guard > 1000 * n + 1000:
            raise RuntimeError("could not place poses inside the level")
        # 1/256-unit grid positions
        px = (x0 << 8) + rng.below(((x1 - x0) << 8) + 1)
        py = (y0 << 8) + rng.below(((y1 - y0) << 8) + 1)
        ang = rng.next() & 0xFFFFFFFF
        sec, fl, ce = scene.sector_at(px / 256.0, py / 256.0)
        if sec < 0 or ce - fl < MIN_ROOM_HEIGHT:
            continue
        out[k] = (px << 8, py << 8, (fl + EYE_HEIGHT) << 16, ang)
        k += 1
    return out


def _catmull_rom(p0, p1, p2, p3, t):
    t2, t3 = t * t, t * t * t
    return 0.5 * ((2 * p1) + (-p0 + p2) * t + (2 * p0 - 5 * p1 + 4 * p2 - p3) * t2 + (-p0 + 3 * p1 - 3 * p2 + p3) * t3)


def flythrough_poses(scene: Scene, n: int, seed: int = 2, waypoints: int = 0) -> np.ndarray:
    verts, ssec, segs = _blob_arrays(scene.blob)
    nss = len(ssec)
    cent = np.zeros((nss, 2))
    edge_owner = {}
    for i in range(nss):
        first, num = int(ssec[i, 0]), int(ssec[i, 1])
        if num == 0 or ssec[i, 2] < 0:
            continue
        pts = []
        for s in segs[first:first + num]:
            if s[3] & 0x80:
                continue
            pts.append(verts[s[0]])
            pts.append(verts[s[1]])
            edge_owner[(int(s[0]), int(s[1]))] = i
        if pts:
            cent[i] = np.mean(np.array(pts, dtype=np.float64), axis=0)
    adj: List[List[int]] = [[] for _ in range(nss)]
    for (a, b), i in edge_owner.items():
        j = edge_owner.get((b, a))
        if j is not None and j != i:
            adj[i].append(j)
    for lst in adj:
        lst.sort()
    rng = SplitMix64(seed)
    start = None
    sp = scene.start_pose
    candidates = [i for i in range(nss) if adj[i]]
    if not candidates:
        raise RuntimeError("level has no connected subsectors")
    if sp is not None:
        sx, sy = sp["x"][0] / 65536.0, sp["y"][0] / 65536.0
        start = min(candidates, key=lambda i: (cent[i, 0] - sx) ** 2 + (cent[i, 1] - sy) ** 2)
    else:
        start = candidates[0]
    nway = waypoints or max(8, n // 12)
    path = [start]
    prev = -1
    while len(path) < nway:
        cur = path[-1]
        opts = [j for j in adj[cur] if j != prev] or adj[cur]
        nxt = opts[rng.below(len(opts))]
        prev = cur
        path.append(nxt)
    P = cent[path]
    P = np.vstack([P[0], P, P[-1]])
    out = np.zeros(n, dtype=POSE_DTYPE)
    nseg = len(path) - 1
    last_floor = None
    for k in range(n):
        u = (k + 0.5) / n * nseg
        i = min(int(u), nseg - 1)
        t = u - i
        p = _catmull_rom(P[i], P[i + 1], P[i + 2], P[i + 3], t)
        q = _catmull_rom(P[i], P[i + 1], P[i + 2], P[i + 3], min(t + 1e-3, 1.0))
        d = q - p
        if abs(d[0]) + abs(d[1]) < 1e-9:
            d = P[i + 2] - P[i + 1]
        yaw = math.atan2(d[1], d[0])
        sec, fl, ce = scene.sector_at(float(p[0]), float(p[1]))
        if sec < 0 or ce - fl < MIN_ROOM_HEIGHT:
            # spline bulged into rock / a pillar: fall back to the straight chord
            p = P[i + 1] + (P[i + 2] - P[i + 1]) * t
            sec, fl, ce = scene.sector_at(float(p[0]), float(p[1]))
        if sec >= 0:
            last_floor = fl
        if last_floor is None:
            last_floor = 0
        out[k] = (int(round(p[0] * 65536.0)), int(round(p[1] * 65536.0)), (last_floor + EYE_HEIGHT) << 16,
                  int(round(yaw / (2 * math.pi) * 4294967296.0)) & 0xFFFFFFFF)
    return out
