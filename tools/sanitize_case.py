#!/usr/bin/env python3
"""A small render through every kernel variant, meant to run under compute-sanitizer (memcheck / racecheck):
plain level (index + RGBA), content-rich level (masked middles, sprites, animation, set_time), a 1-rank sharded render
with the checksum consumer.  Frames are compared with the oracle, so the run also fails on wrong pixels.
usage: compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import rust_doom_b200 as b2d  # noqa: E402
from oracle import render  # noqa: E402
from rust_doom_b200 import jobs, synthwad  # noqa: E402
from rust_doom_b200 import poses as P  # noqa: E402


def main():
    for cfg, tics in ((synthwad.SynthConfig(), 0), (synthwad.SynthConfig(mid_pct=35, thing_pct=50, anim=True, odd_tex=True), 77)):
        sc = b2d.Scene(b2d.Archive.from_bytes(synthwad.build_iwad(3, ("E1M1",), cfg=cfg)), 0)
        poses = np.concatenate([sc.start_pose, P.random_poses(sc, 5, 9)])
        for (w, h) in ((320, 200), (1920, 64), (333, 97)):
            r = b2d.Renderer(sc, b2d.make_view(w, h), device=0, max_batch=4)
            r.set_time(tics)
            idx, rgba = r.render(poses, rgba=True)
            ofb, orgba = render.render(sc.blob, render.make_view(w, h), poses, rgba=True, tics=tics)
            assert np.array_equal(idx, ofb) and np.array_equal(rgba, orgba), (w, h)
            assert r.status() == 0
    comm = jobs.single_comm(0)
    r = b2d.Renderer(sc, b2d.make_view(320, 200), device=0, max_batch=3)
    table = jobs.ChecksumTable(1, len(poses), 320 * 200, torch.device("cuda", 0))
    r.set_time(0)
    st = r.render_sharded(comm, poses, 3, b2d._lib.SHARD_RENDER_GATHER, table.on_chunk)
    ofb = render.render(sc.blob, render.make_view(320, 200), poses)
    assert table.host()[0].tolist() == [b2d.frame_checksum(f) for f in ofb], st
    comm.close()
    print("sanitize case ok")
    return 0


if __name__ == "__main__":
    sys.exit(main())
