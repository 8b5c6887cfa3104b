"""Scene-semantics pin for the oracle.  The reference has no CPU renderer to compare pixels with, but its
pipeline is well defined: a static triangle soup + depth test + the GLSL in assets/shaders/.  With a depth
test the visible surface at a pixel is the nearest one along the pixel's ray, so tests/refcheck/glcaster.py
ray-casts the level in float64 straight from the reference's geometry and shader rules (no BSP order, no column
clipping, no fixed point).  The oracle -- a completely different algorithm -- must agree with it on almost every
pixel; the residue is float-vs-fixed-point rounding at texel / colormap-row / silhouette boundaries."""
import numpy as np
import pytest

from oracle import render, scene, wad
from tests.refcheck import glcaster


def _poses_in(level, want_sky, n, seed):
    rng = np.random.default_rng(seed)
    sec_bytes = level.sectors.tobytes()
    out = []
    while len(out) < n:
        x, y = rng.uniform(-1280, 1280), rng.uniform(-1152, 1152)
        sec = scene.sector_at(level, x, y)
        if sec < 0:
            continue
        s = level.sectors[sec]
        is_sky = wad.is_sky_flat(wad.wad_name(sec_bytes[sec * 26 + 12:sec * 26 + 20]))
        if is_sky != want_sky or int(s["ceil"]) - int(s["floor"]) < 56:
            continue
        out.append((round(x * 4) / 4, round(y * 4) / 4, int(s["floor"]) + 41, float(rng.integers(0, 720)) / 2))
    return out


@pytest.mark.parametrize("want_sky", [False, True])
def test_oracle_agrees_with_reference_semantics_raycaster(synth_wad, oracle_scene, want_sky):
    a = wad.Archive(synth_wad)
    tex = wad.TextureDirectory(a)
    level = wad.Level(a, 0)
    W_, H_ = 320, 200
    view = render.make_view(W_, H_)
    fracs, sky_share = [], []
    for (x, y, z, ang) in _poses_in(level, want_sky, 5, 7 + want_sky):
        g, kind = glcaster.render(a, tex, 0, W_, H_, x, y, z, ang, focal2=(view.F, view.FY2))
        o = render.render(oracle_scene, view, render.make_pose(x, y, z, ang))[0]
        fracs.append(float((g == o).mean()))
        sky_share.append(float((kind == 3).mean()))
        if (kind == 3).any():
            assert (g == o)[kind == 3].mean() > 0.97, "sky mapping disagrees"
    assert min(fracs) > 0.985, fracs
    assert float(np.mean(fracs)) > 0.992, fracs
    if want_sky:
        assert max(sky_share) > 0.05, "no pose actually saw the sky"


def test_masked_middle_textures_agree_with_raycaster():
    """Two-sided middle textures with holes (visitor.rs:808-836; transparent texels discarded,
    static.frag:21-22): the oracle's deferred back-to-front pass vs rays that pass through the holes."""
    from rust_doom_b200 import synthwad
    data = synthwad.build_iwad(1, ("E1M1",), cfg=synthwad.SynthConfig(mid_pct=45))
    a = wad.Archive(data)
    tex = wad.TextureDirectory(a)
    blob = scene.compile_scene(a, tex, 0)
    level = wad.Level(a, 0)
    assert scene.header(blob)[scene.H_NMIDS] > 50
    view = render.make_view(320, 200)
    fracs, twice = [], 0
    for (x, y, z, ang) in _poses_in(level, False, 6, 17) + _poses_in(level, True, 4, 18):
        g, _ = glcaster.render(a, tex, 0, 320, 200, x, y, z, ang, focal2=(view.F, view.FY2))
        o, hits = render.render(blob, view, render.make_pose(x, y, z, ang), seg_hits=True)
        fracs.append(float((g == o[0]).mean()))
        twice += int(hits.sum()) - 320 * 200          # pixels overdrawn by masked textures
    assert min(fracs) > 0.985 and float(np.mean(fracs)) > 0.992, fracs
    assert twice > 5000, "the poses never looked through a masked texture"


def test_decoration_sprites_agree_with_raycaster():
    """Thing sprites (visitor.rs:1062-1137, sprite.vert/frag): the oracle's per-subsector deferred billboards vs
    depth-tested billboards in the ray caster."""
    from rust_doom_b200 import synthwad
    data = synthwad.build_iwad(1, ("E1M1",), cfg=synthwad.SynthConfig(mid_pct=30, thing_pct=70))
    a = wad.Archive(data)
    tex = wad.TextureDirectory(a)
    blob = scene.compile_scene(a, tex, 0)
    level = wad.Level(a, 0)
    assert scene.header(blob)[scene.H_NSPRITES] > 30
    view = render.make_view(320, 200)
    fracs, sprite_px, sprite_same = [], 0, 0
    for (x, y, z, ang) in _poses_in(level, False, 6, 27) + _poses_in(level, True, 4, 28):
        g, kind = glcaster.render(a, tex, 0, 320, 200, x, y, z, ang, focal2=(view.F, view.FY2))
        o = render.render(blob, view, render.make_pose(x, y, z, ang))[0]
        fracs.append(float((g == o).mean()))
        sprite_px += int((kind == 4).sum())
        sprite_same += int(((g == o) & (kind == 4)).sum())
    assert min(fracs) > 0.985 and float(np.mean(fracs)) > 0.992, fracs
    assert sprite_px > 2000 and sprite_same / sprite_px > 0.97, (sprite_px, sprite_same)


def test_animation_and_scrolling_agree_with_raycaster():
    """u_time semantics (static.vert:23-39, visitor.rs:922): frame = floor(tics/8) mod n, scroll = 1 texel per tic."""
    from rust_doom_b200 import synthwad
    data = synthwad.build_iwad(1, ("E1M1",), cfg=synthwad.SynthConfig(anim=True))
    a = wad.Archive(data)
    tex = wad.TextureDirectory(a)
    blob = scene.compile_scene(a, tex, 0)
    level = wad.Level(a, 0)
    view = render.make_view(320, 200)
    fracs, moved = [], 0
    poses = _poses_in(level, False, 4, 37) + _poses_in(level, True, 2, 38)
    for tics in (0, 5, 8, 19, 1000):
        for (x, y, z, ang) in poses:
            g, _ = glcaster.render(a, tex, 0, 320, 200, x, y, z, ang, focal2=(view.F, view.FY2), tics=tics)
            o = render.render(blob, view, render.make_pose(x, y, z, ang), tics=tics)[0]
            o0 = render.render(blob, view, render.make_pose(x, y, z, ang))[0]
            fracs.append(float((g == o).mean()))
            moved += int((o != o0).sum())
    assert min(fracs) > 0.985 and float(np.mean(fracs)) > 0.992, fracs
    assert moved > 20000, "time never changed a pixel"
