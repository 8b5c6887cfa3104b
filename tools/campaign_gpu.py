#!/usr/bin/env python3
"""Random parity campaign on the GPU: libb2d.so (CUDA kernels through the C ABI) against the oracle on random
generated levels (all content kinds), random resolutions, fields of view, level times and (every third case) moved sectors.  Same generator as
tools/campaign.py.  usage: python tools/campaign_gpu.py [cases]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import rust_doom_b200 as b2d  # noqa: E402
from oracle import render, scene as oscene, wad as owad  # noqa: E402
from rust_doom_b200 import synthwad  # noqa: E402
from tests.conftest import sample_poses  # noqa: E402
from tests.refcheck import moves as MV  # noqa: E402


def main(cases=None, seed=12345):
    if cases is None:
        cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rng = np.random.default_rng(seed)
    bad, t0, pixels = 0, time.time(), 0
    for it in range(cases):
        seed = int(rng.integers(100, 100000))
        cfg = synthwad.SynthConfig(mid_pct=int(rng.integers(0, 50)), thing_pct=int(rng.integers(0, 60)),
                                   anim=bool(rng.integers(0, 2)), odd_tex=bool(rng.integers(0, 2)),
                                   rock_pct=int(rng.integers(5, 30)), sky_pct=int(rng.integers(0, 40)),
                                   door_pct=int(rng.integers(5, 40)))
        name = ["E1M1", "E2M3", "MAP05", "MAP15", "MAP25"][int(rng.integers(0, 5))]
        data = synthwad.build_iwad(seed, (name,), cfg=cfg)
        dyn, moves, oblob = (), (), None
        if it % 3 == 2:                                 # every third case: some sectors declared dynamic and moved (C16)
            oa = owad.Archive(data)
            level = owad.Level(oa, 0)
            dyn = MV.declare(level, seed, 10)
            moves = MV.state(level, dyn, seed + 1, hole_free=False)
            oblob = oscene.apply_moves(oscene.compile_scene(oa, owad.TextureDirectory(oa), 0, dynamic=dyn), moves)   # the oracle's own scene
        sc = b2d.Scene(b2d.Archive.from_bytes(data), 0, dynamic=dyn)
        w, h = int(rng.integers(40, 1300)), int(rng.integers(30, 900))
        tics = int(rng.integers(0, 1 << 32)) if rng.integers(0, 2) else 0
        fov = float(rng.uniform(40, 110))
        poses = sample_poses(b2d, sc, 4, seed)
        want_rgba = it % 4 == 3                         # every fourth case also materialises RGBA8 frames
        o = render.render(oblob if oblob is not None else sc.blob, render.make_view(w, h, fov), poses, threads=8, tics=tics, rgba=want_rgba)
        r = b2d.Renderer(sc, b2d.make_view(w, h, fov), max_batch=int(rng.integers(1, 5)))
        r.set_time(tics)
        r.set_sector_moves(moves)
        g = r.render(poses, rgba=want_rgba)
        if want_rgba:
            if not np.array_equal(o[1], g[1]):
                bad += 1
                print("RGBA MISMATCH", seed, cfg, name, w, h, tics, fov, moves)
            o, g = o[0], g[0]
        pixels += o.size
        if not np.array_equal(o, g):
            bad += 1
            print("MISMATCH", seed, cfg, name, w, h, tics, fov, moves, int((o != g).sum()))
    print("gpu campaign: %d cases, %d mismatching, %.1f Mpixel compared, %.1f s" % (cases, bad, pixels / 1e6, time.time() - t0))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
