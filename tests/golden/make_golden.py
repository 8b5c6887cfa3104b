"""Regenerates tests/golden/frames.json: CRC-32 of the ORACLE's palette-index framebuffer for fixed integer
poses on seeded synthetic IWADs.  The reference has no golden frames, so these pin our own
pixel contract: any change to the oracle or the WAD generator shows up here.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from oracle import render, scene, wad  # noqa: E402
from rust_doom_b200 import synthwad  # noqa: E402

CASES = [
    dict(name="e1m1_320x200", seed=1, maps=["E1M1"], level=0, w=320, h=200, n=12),
    dict(name="e1m1_1920x1080", seed=1, maps=["E1M1"], level=0, w=1920, h=1080, n=3),
    dict(name="map12_640x400", seed=21, maps=["MAP01", "MAP12"], level=1, w=640, h=400, n=6),
    dict(name="e1m1_3840x2160", seed=1, maps=["E1M1"], level=0, w=3840, h=2160, n=1),
    # masked middles + decoration sprites + animated / scrolling walls at level time 100 tics
    dict(name="e1m1_full_640x400_t100", seed=5, maps=["E1M1"], level=0, w=640, h=400, n=8, tics=100,
         cfg=dict(mid_pct=30, thing_pct=50, anim=True)),
]


def golden_poses(blob: bytes, n: int, salt: int) -> np.ndarray:
    """Integer poses derived from the scene only: spawn + points stepped around it."""
    h = scene.header(blob)
    s32 = lambda v: v - (1 << 32) if v >= (1 << 31) else v  # noqa: E731
    sx, sy, sz, ang = s32(h[scene.H_START_X]), s32(h[scene.H_START_Y]), s32(h[scene.H_START_Z]), h[scene.H_START_ANGLE]
    poses = np.zeros(n, dtype=render.POSE)
    rng = synthwad.SplitMix64(1000 + salt)
    for k in range(n):
        poses[k] = ((sx << 16) + (rng.below(1 << 22) - (1 << 21)) * (k > 0),
                    (sy << 16) + (rng.below(1 << 22) - (1 << 21)) * (k > 0),
                    (sz - 21 + (rng.below(17) if k else 21)) << 16,
                    (ang * ((1 << 32) // 360) + (rng.next() & 0xFFFFFFFF) * (k > 0)) & 0xFFFFFFFF)
    return poses


def main():
    out = []
    for c in CASES:
        data = synthwad.build_iwad(c["seed"], c["maps"], cfg=synthwad.SynthConfig(**c.get("cfg", {})))
        a = wad.Archive(data)
        blob = scene.compile_scene(a, wad.TextureDirectory(a), c["level"])
        poses = golden_poses(blob, c["n"], c["seed"])
        fb = render.render(blob, render.make_view(c["w"], c["h"]), poses, threads=8, tics=c.get("tics", 0))
        out.append(dict(c, wad_crc=render.crc32(np.frombuffer(data, np.uint8)),
                        blob_crc=render.crc32(np.frombuffer(blob, np.uint8)),
                        poses=[[int(v) for v in p] for p in poses.tolist()],
                        frame_crc=[render.crc32(fb[i]) for i in range(len(poses))]))
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "frames.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    main()
