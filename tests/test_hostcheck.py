"""The product's device maths (b2d_math.cuh) and the kernels' exact algorithms, executed on the CPU by
tests/hostcheck, must reproduce the oracle's framebuffer bit for bit.  This is the no-GPU gate for the
pixel contract; the -m gpu tests repeat it through the real kernels."""
import numpy as np
import pytest

from oracle import render
from tests.conftest import sample_poses


def _compare(b2d, hostcheck, scene, w, h, n, seed, tics=0):
    blob = scene.blob
    poses = sample_poses(b2d, scene, n, seed)
    ofb, hits = render.render(blob, render.make_view(w, h), poses, threads=4, seg_hits=True, tics=tics)
    hfb, counts, ids = hostcheck(blob, b2d.make_view(w, h), poses, tics=tics)
    bad = [(i, int((ofb[i] != hfb[i]).sum())) for i in range(len(poses)) if not np.array_equal(ofb[i], hfb[i])]
    assert not bad, "frames differ (index, pixels): %s" % bad[:5]
    for i in range(len(poses)):
        assert counts[i] >= 0
        drawn = set(np.nonzero(hits[i] > 0)[0].tolist())
        listed = ids[i, :counts[i]].tolist()
        assert drawn <= set(listed), "culling dropped a seg the oracle drew"
        assert len(listed) == len(set(listed))


def test_hostcheck_320x200(b2d, hostcheck, product_scene):
    _compare(b2d, hostcheck, product_scene, 320, 200, 40, 11)


def test_hostcheck_odd_sizes(b2d, hostcheck, product_scene):
    _compare(b2d, hostcheck, product_scene, 333, 187, 8, 12)      # width not a multiple of 32
    _compare(b2d, hostcheck, product_scene, 64, 48, 8, 13)


def test_hostcheck_1080p_and_4k(b2d, hostcheck, product_scene):
    _compare(b2d, hostcheck, product_scene, 1920, 1080, 3, 14)
    _compare(b2d, hostcheck, product_scene, 3840, 2160, 1, 15)


@pytest.mark.parametrize("seed,maps,level", [(7, ("E2M3",), 0), (21, ("MAP01", "MAP12"), 1)])
def test_hostcheck_other_maps(b2d, hostcheck, seed, maps, level):
    from rust_doom_b200 import synthwad
    a = b2d.Archive.from_bytes(synthwad.build_iwad(seed, maps))
    _compare(b2d, hostcheck, b2d.Scene(a, level), 320, 200, 16, seed)


def test_hostcheck_micro_level_extremes(b2d, hostcheck):
    """Camera almost touching a wall, on a BSP partition line, and far outside the level."""
    from tests.test_scene import _micro_level
    a = b2d.Archive.from_bytes(_micro_level())
    sc = b2d.Scene(a, 0)
    poses = np.concatenate([
        b2d.make_pose(-0.01, 128, 60, 0), b2d.make_pose(0, 128, 60, 180), b2d.make_pose(-255.99, 0.01, 1, 45),
        b2d.make_pose(-128, 128, 127.99, 270), b2d.make_pose(-5000, 9000, 60, 300), b2d.make_pose(128, 128, 30, 123.4),
    ])
    ofb = render.render(sc.blob, render.make_view(320, 200), poses)
    hfb, counts, _ = hostcheck(sc.blob, b2d.make_view(320, 200), poses)
    assert np.array_equal(ofb, hfb)


def test_hostcheck_fuzzed_levels(b2d, hostcheck):
    """Inconsistent maps (random bytes in LINEDEFS..SECTORS that still load): bounding-box / solid-column
    culling must stay conservative because the scene compiler recomputes child boxes from the segs."""
    from oracle import scene as S
    from oracle import wad as W
    from rust_doom_b200 import synthwad
    base = synthwad.build_iwad(1, ("E1M1",), cfg=synthwad.SynthConfig(gx=4, gy=3, origin=(-512, -384)))
    oa0 = W.Archive(base)
    rng = synthwad.SplitMix64(77)
    view, oview = b2d.make_view(160, 100), render.make_view(160, 100)
    poses = np.ascontiguousarray(np.concatenate(
        [b2d.make_pose(-300 + 150 * i, -100 + 60 * i, 30 + 7 * i, 47 * i) for i in range(4)]))
    compared = 0
    for it in range(80):
        data = bytearray(base)
        idx = oa0.levels[0] + 2 + rng.below(7)
        _, pos, size = oa0.lumps[idx]
        for _ in range(1 + rng.below(5)):
            data[pos + rng.below(size)] = rng.below(256)
        try:
            oa = W.Archive(bytes(data))
            ob = S.compile_scene(oa, W.TextureDirectory(oa), 0)
        except W.WadError:
            continue
        ofb = render.render(ob, oview, poses)
        hfb, counts, _ = hostcheck(ob, view, poses)
        if (counts < 0).any():
            continue
        assert np.array_equal(ofb, hfb), "fuzz iteration %d" % it
        compared += 1
    assert compared > 30


def test_hostcheck_masked_middle_textures(b2d, hostcheck):
    """Levels with masked two-sided middle textures (all four pegging variants): deferred back-to-front pass."""
    from rust_doom_b200 import synthwad
    data = synthwad.build_iwad(1, ("E1M1",), cfg=synthwad.SynthConfig(mid_pct=45))
    sc = b2d.Scene(b2d.Archive.from_bytes(data), 0)
    from oracle import scene as S
    assert S.header(sc.blob)[S.H_NMIDS] > 50
    _compare(b2d, hostcheck, sc, 320, 200, 40, 51)
    _compare(b2d, hostcheck, sc, 1920, 1080, 2, 52)


def test_hostcheck_decoration_sprites(b2d, hostcheck):
    """Thing sprites (floor-standing, hanging, rotation-1 fallback, unknown / missing sprites skipped) together
    with masked middle textures."""
    from oracle import scene as S
    from rust_doom_b200 import synthwad
    data = synthwad.build_iwad(1, ("E1M1",), cfg=synthwad.SynthConfig(mid_pct=30, thing_pct=70))
    sc = b2d.Scene(b2d.Archive.from_bytes(data), 0)
    assert S.header(sc.blob)[S.H_NSPRITES] > 30
    _compare(b2d, hostcheck, sc, 320, 200, 48, 71)
    _compare(b2d, hostcheck, sc, 1920, 1080, 2, 72)


def test_hostcheck_animated_and_scrolling(b2d, hostcheck):
    """Level time (C14): animated flats / wall textures and scrolling walls.  The oracle resolves the frame and
    the scroll per drawn column; the product rebuilds three scene tables per time step (scene_at_time)."""
    from oracle import scene as S
    from rust_doom_b200 import synthwad
    data = synthwad.build_iwad(1, ("E1M1",), cfg=synthwad.SynthConfig(mid_pct=20, thing_pct=30, anim=True))
    sc = b2d.Scene(b2d.Archive.from_bytes(data), 0)
    hdr = S.header(sc.blob)
    assert hdr[S.H_NANIM] >= 6
    poses = sample_poses(b2d, sc, 16, 91)
    base = render.render(sc.blob, render.make_view(320, 200), poses, threads=4)
    changed = 0
    for tics in (0, 1, 7, 8, 9, 23, 24, 100, 12345, (1 << 24) - 1, (1 << 24) + 5, 0xFFFFFFFF):
        _compare(b2d, hostcheck, sc, 320, 200, 16, 91, tics=tics)
        changed += int((render.render(sc.blob, render.make_view(320, 200), poses, threads=4, tics=tics) != base).sum())
    assert changed > 20000, "time never changed a pixel"
    _compare(b2d, hostcheck, sc, 1920, 1080, 2, 92, tics=77)


def test_light_effects_product_equals_oracle_over_time(b2d, hostcheck):
    """C15: glow / flash / flicker / strobe light levels in float32 (lights.rs:26-66) — the product's C++ evaluation
    and the oracle's numpy evaluation give the same byte for every effect sector at every sampled tic."""
    import ctypes
    from oracle import scene as S
    from rust_doom_b200 import synthwad
    kinds = set()
    for seed in (1, 2):
        data = synthwad.build_iwad(seed, ("E1M1",), cfg=synthwad.SynthConfig(anim=True))
        blob = b2d.Scene(b2d.Archive.from_bytes(data), 0).blob
        h = S.header(blob)
        n = h[S.H_NSECTORS]
        rec = np.frombuffer(blob, dtype="<u4", count=8 * n, offset=h[S.H_OFF_LIGHTS]).reshape(n, 8)
        kinds |= set(rec[:, 0].tolist())
        buf = (ctypes.c_char * len(blob)).from_buffer_copy(blob)
        rng = np.random.default_rng(seed)
        tics = list(range(0, 300)) + rng.integers(0, 1 << 32, 300, dtype=np.uint64).tolist() + [(1 << 32) - 1, 1 << 24]
        seen = set()
        for t in tics:
            out = np.empty(n, np.int16)
            hostcheck.lib.hostcheck_lights(ctypes.c_void_p(ctypes.addressof(buf)), ctypes.c_uint32(t), ctypes.c_void_p(out.ctypes.data))
            want = S.sector_lights_at(blob, t)
            assert np.array_equal(out, want), (seed, t, out.tolist(), want.tolist())
            seen |= set(map(tuple, np.stack([np.arange(n), out], 1)[out >= 0].tolist()))
        assert len(seen) > 3 * int((rec[:, 0] != 0).sum()), "lights never changed over time"
    assert kinds == {0, 1, 2, 3}, kinds


def test_hostcheck_odd_texture_sizes(b2d, hostcheck):
    """Heights that are not a multiple of 4 and widths that are not a power of two: the row-major layout of the
    pre-lit planes and the magic-reciprocal floor-mod, next to the 4-row interleaved textures of the same level."""
    from oracle import scene as S
    from rust_doom_b200 import synthwad
    data = synthwad.build_iwad(4, ("E1M1",), cfg=synthwad.SynthConfig(odd_tex=True, mid_pct=20))
    sc = b2d.Scene(b2d.Archive.from_bytes(data), 0)
    h = S.header(sc.blob)
    tex = np.frombuffer(sc.blob, dtype="<u4", count=8 * h[S.H_NTEX], offset=h[S.H_OFF_TEX]).reshape(-1, 8)
    assert {70, 33, 126} <= set(tex[:, 2].tolist())
    _compare(b2d, hostcheck, sc, 320, 200, 32, 61)
    _compare(b2d, hostcheck, sc, 1920, 1080, 2, 62)


def test_hostcheck_campaign_mixed_content(b2d, hostcheck):
    """The GPU campaign's levels (all content kinds, odd sizes, level times) through the CPU execution of the product maths."""
    from rust_doom_b200 import synthwad
    sizes = ((640, 400), (500, 350), (333, 777))
    for i, seed in enumerate(range(31, 37)):
        cfg = synthwad.SynthConfig(mid_pct=10 * (i % 4), thing_pct=15 * (i % 3), anim=bool(i & 1), odd_tex=bool(i & 2))
        sc = b2d.Scene(b2d.Archive.from_bytes(synthwad.build_iwad(seed, ("MAP07",), cfg=cfg)), 0)
        w, h = sizes[i % 3]
        _compare(b2d, hostcheck, sc, w, h, 6, 300 + seed, tics=(0, 9, 123456)[i % 3])


def test_campaign_tool_smoke(hostcheck):
    """tools/campaign.py (random CPU parity campaign) keeps running: 6 cases, no mismatch."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "campaign.py"), "6"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    assert "bad 0" in out.stdout, out.stdout[-500:]


def test_level_time_is_periodic_without_light_effects(b2d, hostcheck):
    """Size-independent property of C14: with static lights, a frame depends on time only through the animation
    frame (period 8 tics x lcm(2, 3, 4) frames = 96) and the scroll offset (mod the texture widths, all dividing 256),
    so it repeats every lcm(96, 256) = 768 tics -- and does change within the period."""
    from oracle import scene as S
    from rust_doom_b200 import synthwad
    data = synthwad.build_iwad(7, ("E1M1",), cfg=synthwad.SynthConfig(anim=True, light_fx=False, mid_pct=15))
    sc = b2d.Scene(b2d.Archive.from_bytes(data), 0)
    assert (S.sector_lights_at(sc.blob, 0) < 0).all() and S.header(sc.blob)[S.H_NANIM] >= 6
    poses = sample_poses(b2d, sc, 10, 77)
    ov, pv = render.make_view(480, 270), b2d.make_view(480, 270)
    for t in (0, 5, 1234567):
        a = render.render(sc.blob, ov, poses, threads=4, tics=t)
        b = render.render(sc.blob, ov, poses, threads=4, tics=t + 768)
        c = render.render(sc.blob, ov, poses, threads=4, tics=t + 8)
        assert np.array_equal(a, b), "oracle not periodic at t=%d" % t
        assert not np.array_equal(a, c), "time had no effect"
        ha, hb = hostcheck(sc.blob, pv, poses, tics=t)[0], hostcheck(sc.blob, pv, poses, tics=t + 768)[0]
        assert np.array_equal(ha, a) and np.array_equal(hb, a)


def test_every_frame_name_of_a_group_shows_the_same_image(b2d, hostcheck):
    """tex.rs:260 / 302-306: every frame name of an animation group is bound to the atlas position of the group's first
    frame, so NUKAGE3 / NUKAGE2 floors and SFALL2 / SFALL4 / FIREBLU2 walls render exactly like NUKAGE1 / SFALL1 /
    FIREBLU1 at every tic (frame floor(tics/8) mod n of the group; frame 0 at tic 0).  Rewriting the k > 0 names of a
    generated level to the k = 0 name must not change a pixel, for the oracle and for the product's tables."""
    from oracle import scene as S, wad as W
    from rust_doom_b200 import synthwad
    data = synthwad.build_iwad(3, ("E1M1",), cfg=synthwad.SynthConfig(anim=True, light_fx=False))
    same = bytearray(data)
    renames = {b"NUKAGE3\0": b"NUKAGE1\0", b"NUKAGE2\0": b"NUKAGE1\0", b"SFALL2\0\0": b"SFALL1\0\0", b"SFALL4\0\0": b"SFALL1\0\0",
               b"FIREBLU2": b"FIREBLU1"}
    arch = W.Archive(data)
    hits = 0
    marker = arch.levels[0]
    for lump in (3, 8):                             # SIDEDEFS, SECTORS of level 0 (wad/src/level.rs:13-20): only the
        _, off, size = arch.lumps[marker + lump]     # level's references, not the texture / flat definitions
        body = bytes(same[off:off + size])
        for old, new in renames.items():
            hits += body.count(old)
            body = body.replace(old, new)
        same[off:off + size] = body
    assert hits >= 4, "the generated level uses no k > 0 frame names"
    sc_k = b2d.Scene(b2d.Archive.from_bytes(data), 0)
    sc_0 = b2d.Scene(b2d.Archive.from_bytes(bytes(same)), 0)
    poses = sample_poses(b2d, sc_k, 12, 5)
    v = render.make_view(320, 200)
    pv = b2d.make_view(320, 200)
    for tics in (0, 7, 8, 16, 31, 1000):
        a = render.render(sc_k.blob, v, poses, threads=4, tics=tics)
        b = render.render(sc_0.blob, v, poses, threads=4, tics=tics)
        assert np.array_equal(a, b), "oracle: a k > 0 frame name rendered differently at tic %d" % tics
        ha, _, _ = hostcheck(sc_k.blob, pv, poses, tics=tics)
        assert np.array_equal(ha, a), "product tables differ from the oracle at tic %d" % tics


def test_hostcheck_moving_sectors(b2d, hostcheck):
    """Doors / lifts: the product's re-derivation of the height-dependent tables (scene_at_time with the
    sector offsets) under the kernels' algorithms, against the oracle rendering its own scene with the same moves applied
    by the numpy restatement (oracle/scene.py apply_moves) -- including closed doors (ceiling on the floor), states the
    reference itself cannot show without holes, level time on top, and eye points inside moved sectors."""
    from oracle import scene as S, wad as W
    from rust_doom_b200 import synthwad
    from tests.refcheck import moves as MV
    data = synthwad.build_iwad(2, ("E1M1",), cfg=synthwad.SynthConfig(mid_pct=30, thing_pct=50, anim=True))
    a = W.Archive(data)
    tex = W.TextureDirectory(a)
    level = W.Level(a, 0)
    dyn = MV.declare(level, 3, 20)
    # a door: one more sector whose ceiling can come all the way down to its floor
    door = next(i for i in range(len(level.sectors)) if i not in {d[0] for d in dyn}
                and int(level.sectors[i]["ceil"]) - int(level.sectors[i]["floor"]) >= 64)
    f0, c0 = int(level.sectors[door]["floor"]), int(level.sectors[door]["ceil"])
    dyn.append((door, f0, f0, f0, c0))
    oblob = S.compile_scene(a, tex, 0, dynamic=dyn)
    sc = b2d.Scene(b2d.Archive.from_bytes(data), 0, dynamic=dyn)
    assert sc.blob == oblob
    poses = sample_poses(b2d, sc, 24, 5)
    view, oview = b2d.make_view(320, 200), render.make_view(320, 200)
    rest = render.render(oblob, oview, poses, threads=4)
    changed = 0
    for k, tics in enumerate((0, 0, 9, 1000)):
        moves = MV.state(level, dyn[:-1], 40 + k, hole_free=False) + [(door, 0, -(c0 - f0) if k % 2 == 0 else -((c0 - f0) // 2))]
        ofb = render.render(S.apply_moves(oblob, moves), oview, poses, threads=4, tics=tics)
        hfb, counts, _ = hostcheck(oblob, view, poses, tics=tics, moves=moves)
        bad = [(i, int((ofb[i] != hfb[i]).sum())) for i in range(len(poses)) if not np.array_equal(ofb[i], hfb[i])]
        assert not bad, "state %d: frames differ (index, pixels): %s" % (k, bad[:5])
        assert (counts >= 0).all()
        changed += int((ofb != rest).sum())
    assert changed > 100000, "the moves never changed a pixel"
    hfb, _, _ = hostcheck(oblob, view, poses, moves=())
    assert np.array_equal(hfb, rest)
    mv1080 = MV.state(level, dyn[:-1], 50, hole_free=False)
    o = render.render(S.apply_moves(oblob, mv1080), render.make_view(1920, 1080), poses[:2], threads=2)
    hfb, _, _ = hostcheck(oblob, b2d.make_view(1920, 1080), poses[:2], moves=mv1080)
    assert np.array_equal(o, hfb)
