"""Synthetic IWAD generator (test fixture *and* benchmark input).

No WAD file ships with the reference (its .gitignore excludes ``*.wad``) and none exists in
the build or GPU environment, so every test and measurement runs on IWADs produced here.  The
output is a byte-valid IWAD in the layout the reference's loader expects
(reference: wad/src/archive.rs:62-106, wad/src/types.rs:20-169, wad/src/level.rs:13-20):
PLAYPAL (14 palettes), COLORMAP (34 maps), PNAMES + patch lumps in Doom picture format,
TEXTURE1 (multi-patch composites, negative origins, a non-power-of-two height), F_START..F_END
flats incl. F_SKY1, S_START..S_END, and per level the lumps THINGS, LINEDEFS, SIDEDEFS,
VERTEXES, SEGS, SSECTORS, NODES, SECTORS, REJECT, BLOCKMAP with a *correct* BSP.

Level shape: a grid of square cells.  A cell is either solid rock or a room (one sector).
Room/room edges are open two-sided lines, doorways (solid-open-solid) or zero-thickness solid
walls; room/rock edges are one-sided.  Some rooms hold an inner convex polygon that is either a
solid pillar or a raised/lowered platform sector, which gives diagonal walls, split segs with
non-zero ``offset`` and BSP sub-trees whose partition lines are not axis aligned.  The BSP is a
k-d split along grid lines (never cuts a seg) followed by per-cell cuts along the inner polygon's
edges (seg split points are integer by construction; asserted).

Everything is driven by a splitmix64 stream so output bytes depend only on the arguments.
"""
from __future__ import annotations

import math
import struct
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

__all__ = ["build_iwad", "SynthConfig", "E1_MAPS", "MAP_NAMES_DOOM2"]

E1_MAPS = tuple("E1M%d" % i for i in range(1, 10))
MAP_NAMES_DOOM2 = tuple("MAP%02d" % i for i in range(1, 11))


# --------------------------------------------------------------------------------------
# deterministic RNG
# --------------------------------------------------------------------------------------
class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def below(self, n: int) -> int:
        return self.next() % n

    def chance(self, num: int, den: int) -> bool:
        return self.below(den) < num

    def pick(self, seq):
        return seq[self.below(len(seq))]

    def bytes2d(self, h: int, w: int) -> np.ndarray:
        n = h * w
        out = np.empty(n, dtype=np.uint8)
        i = 0
        while i < n:
            v = self.next()
            for k in range(8):
                if i < n:
                    out[i] = (v >> (8 * k)) & 0xFF
                    i += 1
        return out.reshape(h, w)


def _iround_hypot(dx: int, dy: int) -> int:
    """round(sqrt(dx^2+dy^2)) in exact integer arithmetic."""
    n = dx * dx + dy * dy
    r = math.isqrt(n)
    return r + (1 if n - r * r > r else 0)


def _name8(s: str) -> bytes:
    b = s.encode("ascii")
    assert len(b) <= 8, s
    return b.ljust(8, b"\0")


# --------------------------------------------------------------------------------------
# palette / colormap
# --------------------------------------------------------------------------------------
_HUES = [
    (255, 255, 255), (200, 160, 120), (255, 64, 48), (96, 220, 96),
    (80, 120, 255), (220, 200, 140), (160, 110, 70), (255, 200, 64),
    (140, 150, 170), (200, 90, 200), (90, 200, 210), (120, 96, 80),
    (255, 140, 60), (170, 200, 110), (110, 100, 160), (230, 120, 130),
]


def make_playpal() -> bytes:
    """14 palettes x 256 RGB.  Palette 0: 16 hues x 16 shades (index = hue*16 + shade, shade 0
    brightest).  Palettes 1..13: red/yellow/green tints of palette 0 (never used by the path)."""
    base = np.zeros((256, 3), dtype=np.int32)
    for hue, rgb in enumerate(_HUES):
        for shade in range(16):
            k = 255 - shade * 15
            base[hue * 16 + shade] = [(c * k) // 255 for c in rgb]
    pals = [base]
    for p in range(1, 14):
        tint = np.array([(255, 0, 0), (215, 186, 69), (0, 255, 0)][(p - 1) % 3], dtype=np.int32)
        a = 1 + (p - 1) // 3
        pals.append((base * (8 - a) + tint * a) // 8)
    return np.stack(pals).astype(np.uint8).tobytes()


def make_colormap(playpal: bytes) -> bytes:
    """34 maps x 256.  Map k<32 sends colour i to the palette entry nearest to rgb(i)*(32-k)/32;
    map 32 = inverse greyscale (invulnerability); map 33 = black."""
    pal = np.frombuffer(playpal[:768], dtype=np.uint8).reshape(256, 3).astype(np.int64)

    def nearest(target: np.ndarray) -> np.ndarray:
        d = ((target[:, None, :] - pal[None, :, :]) ** 2).sum(axis=2)
        return d.argmin(axis=1).astype(np.uint8)

    # darkening stays inside the colour's own 16-shade ramp (index = hue*16 + shade,
    # brightness 255 - 15*shade), so light diminishing never changes hue
    maps = []
    idx = np.arange(256)
    hue, shade = idx // 16, idx % 16
    for k in range(32):
        target = (255 - 15 * shade) * (32 - k)              # brightness * 32
        cand = (255 - 15 * np.arange(16)) * 32
        best = np.abs(target[:, None] - cand[None, :]).argmin(axis=1)
        maps.append((hue * 16 + best).astype(np.uint8))
    lum = (pal[:, 0] * 299 + pal[:, 1] * 587 + pal[:, 2] * 114) // 1000
    inv = np.stack([255 - lum] * 3, axis=1)
    maps.append(nearest(inv))
    maps.append(nearest(np.zeros_like(pal)))
    return np.stack(maps).tobytes()


# --------------------------------------------------------------------------------------
# procedural images -> Doom picture format
# --------------------------------------------------------------------------------------
def _img_bricks(rng: SplitMix64, w: int, h: int, hue: int, bw: int, bh: int) -> np.ndarray:
    noise = rng.bytes2d(h, w)
    ys, xs = np.mgrid[0:h, 0:w]
    row = ys // bh
    xoff = (row % 2) * (bw // 2)
    mortar = ((ys % bh) == 0) | (((xs + xoff) % bw) == 0)
    shade = 3 + (noise % 4) + ((row * 5 + (xs + xoff) // bw * 3) % 4)
    shade = np.where(mortar, 12 + noise % 3, shade)
    return (hue * 16 + np.clip(shade, 0, 15)).astype(np.int16)


def _img_panels(rng: SplitMix64, w: int, h: int, hue: int, hue2: int) -> np.ndarray:
    noise = rng.bytes2d(h, w)
    ys, xs = np.mgrid[0:h, 0:w]
    shade = 2 + (ys * 10) // max(h, 1) + noise % 3
    img = hue * 16 + np.clip(shade, 0, 15)
    stripe = ((xs // 8 + ys // 16) % 4) == 0
    img = np.where(stripe, hue2 * 16 + np.clip(4 + noise % 5, 0, 15), img)
    rivet = ((xs % 16) == 3) & ((ys % 16) == 3)
    img = np.where(rivet, 0 * 16 + 1, img)
    return img.astype(np.int16)


def _img_gradient(rng: SplitMix64, w: int, h: int, hue: int) -> np.ndarray:
    noise = rng.bytes2d(h, w)
    ys, xs = np.mgrid[0:h, 0:w]
    shade = ((xs * 7 + ys * 3) // 9 + noise % 2) % 16
    return (hue * 16 + shade).astype(np.int16)


def _img_grate(rng: SplitMix64, w: int, h: int, hue: int) -> np.ndarray:
    """Overlay with transparent holes (-1)."""
    noise = rng.bytes2d(h, w)
    ys, xs = np.mgrid[0:h, 0:w]
    bars = ((xs % 8) < 2) | ((ys % 12) < 2)
    img = np.where(bars, (hue * 16 + 5 + noise % 4).astype(np.int16), np.int16(-1))
    return img.astype(np.int16)


def _img_sprite(rng: SplitMix64, w: int, h: int, hue: int) -> np.ndarray:
    """Elliptic blob with transparent surround (-1), shaded top to bottom."""
    noise = rng.bytes2d(h, w)
    ys, xs = np.mgrid[0:h, 0:w]
    inside = ((2 * xs + 1 - w) * h) ** 2 + ((2 * ys + 1 - h) * w) ** 2 <= (w * h) ** 2
    shade = np.clip(1 + (ys * 11) // max(h, 1) + (noise % 3).astype(np.int16), 0, 15)
    img = np.where(inside, (hue * 16 + shade).astype(np.int16), np.int16(-1))
    stripe = inside & ((ys % 9) == 4)
    return np.where(stripe, np.int16(7 * 16 + 3), img).astype(np.int16)


def _tri(v: np.ndarray, period: int, amp: int) -> np.ndarray:
    """Integer triangle wave in [-amp, amp] (no libm: WAD bytes must not depend on the host)."""
    ph = np.mod(v, period)
    half = period // 2
    t = np.where(ph < half, ph, period - ph)          # 0..half
    return (t * 2 * amp) // half - amp


def _img_sky(rng: SplitMix64, w: int, h: int) -> np.ndarray:
    noise = rng.bytes2d(h, w)
    ys, xs = np.mgrid[0:h, 0:w]
    cloud = (_tri(xs, 56, 2) + _tri(xs + ys * 2, 106, 2) + _tri(ys, 32, 2))
    shade = np.clip(2 + (ys * 9) // h + cloud + noise % 2, 0, 15)
    img = 4 * 16 + shade
    mountain = ys > (h - 30 + _tri(xs, 82, 10) + _tri(xs, 32, 4))
    img = np.where(mountain, 11 * 16 + np.clip(6 + noise % 6, 0, 15), img)
    return img.astype(np.int16)


def encode_picture(img: np.ndarray, xoff: int = 0, yoff: int = 0) -> bytes:
    """Doom picture format (reference decoder: wad/src/image.rs:39-169): 8-byte header, w column
    offsets, columns of posts {topdelta, len, pad, pixels, pad}, 0xFF terminator. -1 = hole."""
    h, w = img.shape
    assert h < 255 and w <= 4096
    cols = []
    for x in range(w):
        col = img[:, x]
        out = bytearray()
        y = 0
        while y < h:
            if col[y] < 0:
                y += 1
                continue
            y0 = y
            while y < h and col[y] >= 0 and (y - y0) < 128:
                y += 1
            run = bytes(int(v) & 0xFF for v in col[y0:y])
            out += bytes([y0, len(run), 0]) + run + b"\0"
        out += b"\xff"
        cols.append(bytes(out))
    header = struct.pack("<HHhh", w, h, xoff, yoff)
    offs = []
    pos = 8 + 4 * w
    for c in cols:
        offs.append(pos)
        pos += len(c)
    return header + struct.pack("<%dI" % w, *offs) + b"".join(cols)


@dataclass
class TexDef:
    name: str
    w: int
    h: int
    patches: List[Tuple[int, int, str]]  # (origin_x, origin_y, patch name)


def make_graphics(rng: SplitMix64, masked: bool = False, anim: bool = False, odd: bool = False):
    """Returns (patch lumps {name: bytes} in order, texture defs, flats {name: 4096 bytes})."""
    patches: Dict[str, bytes] = {}

    def add(name, img):
        patches[name] = encode_picture(img)

    add("WBRICK1", _img_bricks(rng, 64, 128, 2, 32, 16))
    add("WBRICK2", _img_bricks(rng, 64, 128, 6, 16, 8))
    add("WBRICK3", _img_bricks(rng, 128, 128, 1, 32, 16))
    add("WPANEL1", _img_panels(rng, 64, 128, 8, 4))
    add("WPANEL2", _img_panels(rng, 128, 128, 5, 12))
    add("WPANEL3", _img_panels(rng, 64, 72, 10, 7))
    add("WGRAD1", _img_gradient(rng, 64, 128, 3))
    add("WGRAD2", _img_gradient(rng, 32, 64, 9))
    add("WGRAD3", _img_gradient(rng, 128, 64, 13))
    add("WSTEP1", _img_panels(rng, 32, 16, 11, 0))
    add("WSTEP2", _img_bricks(rng, 64, 24, 14, 8, 8))
    add("WGRATE1", _img_grate(rng, 64, 128, 0))
    add("WTECH1", _img_panels(rng, 64, 64, 4, 15))
    add("WTECH2", _img_gradient(rng, 64, 64, 12))
    add("SKY1", _img_sky(rng, 256, 128))

    tex = [
        TexDef("BRICK1", 64, 128, [(0, 0, "WBRICK1")]),
        TexDef("BRICK2", 64, 128, [(0, 0, "WBRICK2")]),
        TexDef("BRICK3", 128, 128, [(0, 0, "WBRICK3")]),
        TexDef("PANEL1", 64, 128, [(0, 0, "WPANEL1")]),
        TexDef("PANEL2", 128, 128, [(0, 0, "WPANEL2")]),
        TexDef("PANEL72", 64, 72, [(0, 0, "WPANEL3")]),          # non power-of-two height
        TexDef("GRAD1", 64, 128, [(0, 0, "WGRAD1")]),
        TexDef("WIDE1", 256, 128, [(0, 0, "WBRICK3"), (128, 0, "WPANEL2")]),
        TexDef("COMBO1", 128, 128, [(0, 0, "WBRICK1"), (64, 0, "WPANEL1"), (32, 32, "WTECH1")]),
        TexDef("COMBO2", 64, 128, [(0, 0, "WGRAD1"), (0, 0, "WGRATE1")]),   # masked overlay
        TexDef("COMBO3", 128, 128, [(0, 0, "WPANEL2"), (-16, -8, "WGRAD2"), (112, 96, "WGRAD2")]),
        TexDef("TECH1", 64, 64, [(0, 0, "WTECH1")]),
        TexDef("TECH2", 128, 64, [(0, 0, "WGRAD3"), (32, 0, "WTECH2")]),
        TexDef("STEP1", 32, 16, [(0, 0, "WSTEP1")]),
        TexDef("STEP2", 64, 24, [(0, 0, "WSTEP2")]),
        TexDef("TALL1", 64, 128, [(0, 0, "WTECH1"), (0, 64, "WTECH2")]),
        TexDef("SKY1", 256, 128, [(0, 0, "SKY1")]),
    ]

    if masked:
        # single-patch textures whose first (opaque-copied) patch has holes -> transparent texels survive
        add("WGRATE2", _img_grate(rng, 32, 64, 7))
        add("WFENCE1", np.where(_img_grate(rng, 64, 72, 3) < 0, -1, _img_bricks(rng, 64, 72, 9, 16, 12)).astype(np.int16))
        tex += [TexDef("GRATE1", 64, 128, [(0, 0, "WGRATE1")]), TexDef("GRATE2", 32, 64, [(0, 0, "WGRATE2")]),
                TexDef("FENCE72", 64, 72, [(0, 0, "WFENCE1")])]

    if anim:
        for k in range(1, 5):
            add("WSFALL%d" % k, _img_bricks(rng, 64, 128, 4, 8, 4 + 2 * k))
            tex.append(TexDef("SFALL%d" % k, 64, 128, [(0, 0, "WSFALL%d" % k)]))
        for k in range(1, 3):
            add("WFIRE%d" % k, _img_gradient(rng, 128, 128, 2 + 10 * (k - 1)))
            tex.append(TexDef("FIREBLU%d" % k, 128, 128, [(0, 0, "WFIRE%d" % k)]))

    if odd:
        # sizes no stock texture has: heights that are not a multiple of 4, widths that are not a power of two
        add("WODD1", _img_bricks(rng, 64, 70, 3, 10, 7))
        add("WODD2", _img_gradient(rng, 48, 33, 11))
        add("WODD3", _img_panels(rng, 100, 126, 7, 5))
        tex += [TexDef("ODD70", 64, 70, [(0, 0, "WODD1")]), TexDef("ODD33", 48, 33, [(0, 0, "WODD2")]),
                TexDef("ODD126", 100, 126, [(0, 0, "WODD3")])]

    flats: Dict[str, bytes] = {}

    def addflat(name, img):
        assert img.shape == (64, 64)
        flats[name] = img.astype(np.uint8).tobytes()

    addflat("FLOOR1", _img_bricks(rng, 64, 64, 6, 16, 16))
    addflat("FLOOR2", _img_bricks(rng, 64, 64, 8, 32, 32))
    addflat("FLOOR3", _img_gradient(rng, 64, 64, 3))
    addflat("FLOOR4", _img_panels(rng, 64, 64, 11, 1))
    addflat("FLOOR5", _img_panels(rng, 64, 64, 5, 2))
    addflat("FLOOR6", _img_gradient(rng, 64, 64, 7))
    addflat("CEIL1", _img_panels(rng, 64, 64, 8, 0))
    addflat("CEIL2", _img_bricks(rng, 64, 64, 1, 8, 8))
    addflat("CEIL3", _img_gradient(rng, 64, 64, 14))
    addflat("CEIL4", _img_panels(rng, 64, 64, 9, 4))
    addflat("NUKAGE1", _img_gradient(rng, 64, 64, 13))
    addflat("F_SKY1", _img_gradient(rng, 64, 64, 4))
    if anim:
        addflat("NUKAGE2", _img_gradient(rng, 64, 64, 3))
        addflat("NUKAGE3", _img_bricks(rng, 64, 64, 13, 8, 8))
    return patches, tex, flats


WALL_TEX = ["BRICK1", "BRICK2", "BRICK3", "PANEL1", "PANEL2", "PANEL72", "GRAD1", "WIDE1",
            "COMBO1", "COMBO2", "COMBO3", "TECH1", "TECH2", "TALL1"]
STEP_TEX = ["STEP1", "STEP2", "TECH1", "BRICK2", "PANEL72"]
FLOOR_FLATS = ["FLOOR1", "FLOOR2", "FLOOR3", "FLOOR4", "FLOOR5", "FLOOR6", "NUKAGE1"]
CEIL_FLATS = ["CEIL1", "CEIL2", "CEIL3", "CEIL4", "FLOOR3"]
MASKED_TEX = ["GRATE1", "GRATE2", "FENCE72"]


def make_pnames_texture1(patch_names: Sequence[str], tex: Sequence[TexDef]) -> Tuple[bytes, bytes]:
    pn = struct.pack("<I", len(patch_names)) + b"".join(_name8(n) for n in patch_names)
    index = {n: i for i, n in enumerate(patch_names)}
    bodies = []
    for t in tex:
        b = struct.pack("<8sIHHIH", _name8(t.name), 0, t.w, t.h, 0, len(t.patches))
        for ox, oy, pname in t.patches:
            b += struct.pack("<hhHHH", ox, oy, index[pname], 1, 0)
        bodies.append(b)
    off = 4 + 4 * len(tex)
    offs = []
    for b in bodies:
        offs.append(off)
        off += len(b)
    t1 = struct.pack("<I", len(tex)) + struct.pack("<%dI" % len(tex), *offs) + b"".join(bodies)
    return pn, t1


# --------------------------------------------------------------------------------------
# level geometry
# --------------------------------------------------------------------------------------
@dataclass
class Sector:
    floor: int
    ceil: int
    floor_flat: str
    ceil_flat: str
    light: int
    stype: int = 0
    tag: int = 0


@dataclass
class Sidedef:
    xoff: int
    yoff: int
    upper: str
    lower: str
    middle: str
    sector: int


@dataclass
class Linedef:
    v1: int
    v2: int
    flags: int
    special: int
    tag: int
    right: int
    left: int


@dataclass
class Seg:
    v1: int
    v2: int
    linedef: int
    direction: int
    offset: int = 0


@dataclass
class SynthConfig:
    gx: int = 10
    gy: int = 9
    cell: int = 256
    origin: Tuple[int, int] = (-1280, -1152)
    rock_pct: int = 14
    inner_pct: int = 34
    sky_pct: int = 18
    wall_pct: int = 18
    door_pct: int = 22
    mid_pct: int = 0            # % of two-sided lines that carry a masked middle texture (0 keeps legacy bytes)
    thing_pct: int = 0          # % of plain room cells that get decoration things + sprite lumps (0 = legacy)
    anim: bool = False          # animated flats/walls (NUKAGE1-3, SFALL1-4, FIREBLU1-2) + scrolling lines (0x30)
    odd_tex: bool = False       # wall textures with odd heights / non power-of-two widths (ODD70, ODD33, ODD126)
    light_fx: bool = True       # sector light specials (flash, glow, strobes); False = every sector static


# inner convex polygons, CCW, in cell-local coordinates for a 256 cell (scaled by cell/256);
# every edge direction is in {axis, 45deg, 1:2, 2:1} and coordinates are even so that the edge
# lines meet the cell boundary at integer points.
_INNER_SHAPES = [
    [(128, 64), (192, 128), (128, 192), (64, 128)],                                   # diamond
    [(96, 64), (160, 64), (192, 96), (192, 160), (160, 192), (96, 192), (64, 160), (64, 96)],  # octagon
    [(64, 96), (128, 64), (192, 96), (192, 160), (128, 192), (64, 160)],              # hexagon 2:1
    [(80, 80), (176, 80), (176, 176), (80, 176)],                                     # square
    [(128, 48), (208, 208), (48, 208)],                                               # triangle 1:2
]


class LevelBuilder:
    def __init__(self, name: str, seed: int, cfg: SynthConfig):
        self.name = name
        self.cfg = cfg
        self.rng = SplitMix64(seed * 0x1000193 + 0xB2D)
        self.vertices: List[Tuple[int, int]] = []
        self.vindex: Dict[Tuple[int, int], int] = {}
        self.sectors: List[Sector] = []
        self.sidedefs: List[Sidedef] = []
        self.linedefs: List[Linedef] = []
        self.segs: List[Seg] = []
        self.ssectors: List[Tuple[int, int]] = []   # (num, first)
        self.nodes: List[tuple] = []
        self.things: List[tuple] = []

    # -- primitives ---------------------------------------------------------------
    def vert(self, x: int, y: int) -> int:
        assert -32768 <= x <= 32767 and -32768 <= y <= 32767
        key = (int(x), int(y))
        if key not in self.vindex:
            self.vindex[key] = len(self.vertices)
            self.vertices.append(key)
        return self.vindex[key]

    def side(self, sector: int, upper="-", lower="-", middle="-", xoff=0, yoff=0) -> int:
        self.sidedefs.append(Sidedef(xoff, yoff, upper, lower, middle, sector))
        return len(self.sidedefs) - 1

    def _nukage(self, flat: str) -> str:
        """With animation on, a NUKAGE floor uses any frame name of its group (k > 0 too: the reference shows the same
        image for all of them, tex.rs:302-306); chosen by sector index so that no random draw is consumed."""
        if self.cfg.anim and flat == "NUKAGE1":
            return ("NUKAGE1", "NUKAGE3", "NUKAGE2")[len(self.sectors) % 3]
        return flat

    def line(self, v1, v2, right, left=-1, flags=None, special=0) -> int:
        if flags is None:
            flags = 0x0001 if left < 0 else 0x0004
        self.linedefs.append(Linedef(v1, v2, flags, special, 0, right, left))
        return len(self.linedefs) - 1

    # -- generation ---------------------------------------------------------------
    def generate(self):
        cfg, rng = self.cfg, self.rng
        gx, gy, cs = cfg.gx, cfg.gy, cfg.cell
        ox, oy = cfg.origin
        # 1. rock / room layout, keep rooms 4-connected
        room = [[True] * gy for _ in range(gx)]
        order = [(i, j) for i in range(gx) for j in range(gy)]
        for _ in range(gx * gy):
            a, b = rng.below(len(order)), rng.below(len(order))
            order[a], order[b] = order[b], order[a]
        target_rock = (gx * gy * cfg.rock_pct) // 100
        rocks = 0
        for (i, j) in order:
            if rocks >= target_rock:
                break
            room[i][j] = False
            if self._connected(room):
                rocks += 1
            else:
                room[i][j] = True
        self.room = room
        # 2. sectors
        cell_sector = [[-1] * gy for _ in range(gx)]
        heights = [[0] * gy for _ in range(gx)]
        sky_seed = [(rng.below(gx), rng.below(gy)) for _ in range(max(1, gx * gy * cfg.sky_pct // 400))]
        for i in range(gx):
            for j in range(gy):
                if not room[i][j]:
                    continue
                near_sky = any(abs(i - a) + abs(j - b) <= 1 for a, b in sky_seed)
                base = ((i * 3 + j * 5) % 7) * 8 - 24
                fl = base + rng.below(5) * 8 - 16
                if rng.chance(1, 9):
                    fl += rng.pick([-48, 40, 64])
                height = rng.pick([96, 112, 128, 128, 160, 192, 256])
                is_sky = near_sky
                if is_sky:
                    height += 128
                light = rng.pick([96, 112, 128, 144, 160, 176, 192, 208, 224, 255, 255, 80])
                stype = rng.pick([0] * 14 + [1, 8])
                if self.cfg.anim and rng.chance(1, 3):          # every light effect kind (light.rs:127-134)
                    stype = rng.pick([1, 2, 3, 4, 8, 12, 13, 17])
                if not self.cfg.light_fx:
                    stype = 0
                self.sectors.append(Sector(
                    fl, fl + height, self._nukage(rng.pick(FLOOR_FLATS)),
                    "F_SKY1" if is_sky else rng.pick(CEIL_FLATS), light, stype))
                cell_sector[i][j] = len(self.sectors) - 1
                heights[i][j] = fl
        self.cell_sector = cell_sector
        # 3. per-cell seg lists (boundary)
        cell_segs: Dict[Tuple[int, int], List[Seg]] = {(i, j): [] for i in range(gx) for j in range(gy) if room[i][j]}

        def corner(i, j):
            return (ox + i * cs, oy + j * cs)

        def add_seg(cellkey, line_idx, direction, a, b):
            """seg from vertex a to b (seg direction), lying on linedef line_idx."""
            ld = self.linedefs[line_idx]
            anchor = self.vertices[ld.v1] if direction == 0 else self.vertices[ld.v2]
            pa = self.vertices[a]
            off = _iround_hypot(pa[0] - anchor[0], pa[1] - anchor[1])
            cell_segs[cellkey].append(Seg(a, b, line_idx, direction, off))

        def wall_tex():
            if cfg.anim and rng.chance(1, 5):
                return rng.pick(["SFALL2", "SFALL1", "FIREBLU1", "FIREBLU2", "SFALL4"])
            if cfg.odd_tex and rng.chance(1, 3):
                return rng.pick(["ODD70", "ODD33", "ODD126"])
            return rng.pick(WALL_TEX)

        def one_sided(cellkey, p, q, sector):
            """solid wall whose front (right) side faces `sector`; p->q has the sector on the right."""
            a, b = self.vert(*p), self.vert(*q)
            flags = 0x0001 | (0x0010 if rng.chance(1, 4) else 0)
            sd = self.side(sector, middle=wall_tex(), xoff=rng.pick([0, 0, 0, 16, -24, 40]),
                           yoff=rng.pick([0, 0, 0, 8, -16]))
            li = self.line(a, b, sd, -1, flags=flags, special=0x30 if (cfg.anim and rng.chance(1, 6)) else 0)
            add_seg(cellkey, li, 0, a, b)

        def two_sided(keyA, keyB, p, q):
            """open line; cell A is on the right of p->q, cell B on the left."""
            sa, sb = cell_sector[keyA[0]][keyA[1]], cell_sector[keyB[0]][keyB[1]]
            a, b = self.vert(*p), self.vert(*q)
            flags = 0x0004
            if rng.chance(1, 3):
                flags |= 0x0008
            if rng.chance(1, 3):
                flags |= 0x0010
            xo = rng.pick([0, 0, 8, 32])
            yo = rng.pick([0, 0, 0, 4, -12])
            up1, lo1, up2, lo2 = wall_tex(), rng.pick(STEP_TEX + WALL_TEX), wall_tex(), rng.pick(STEP_TEX + WALL_TEX)
            mid = "-"
            if cfg.mid_pct and rng.below(100) < cfg.mid_pct:
                mid = rng.pick(MASKED_TEX)
                style = rng.below(4)          # which of upper/lower are untextured decides the peg (float vs tiled)
                A, B = self.sectors[sa], self.sectors[sb]
                # leave a piece untextured only on a side that does not need it (well-formed map)
                if style in (0, 2):
                    if not (B.ceil < A.ceil):
                        up1 = "-"
                    if not (A.ceil < B.ceil):
                        up2 = "-"
                if style in (1, 2):
                    if not (B.floor > A.floor):
                        lo1 = "-"
                    if not (A.floor > B.floor):
                        lo2 = "-"
            s1 = self.side(sa, upper=up1, lower=lo1, middle=mid, xoff=xo, yoff=yo)
            s2 = self.side(sb, upper=up2, lower=lo2, middle=mid, xoff=xo, yoff=yo)
            li = self.line(a, b, s1, s2, flags=flags)
            add_seg(keyA, li, 0, a, b)
            add_seg(keyB, li, 1, b, a)

        # spanning tree over rooms guarantees reachability through OPEN/DOOR edges
        tree = self._spanning_tree(room)

        def edge(keyA, keyB, p, q):
            """edge p->q with cell A on its right, cell B on its left (either may be rock/None)."""
            ra = keyA is not None and room[keyA[0]][keyA[1]]
            rb = keyB is not None and room[keyB[0]][keyB[1]]
            if ra and rb:
                must_open = (keyA, keyB) in tree or (keyB, keyA) in tree
                r = rng.below(100)
                kind = "open"
                if r < cfg.wall_pct and not must_open:
                    kind = "wall"
                elif r < cfg.wall_pct + cfg.door_pct:
                    kind = "door"
                sa, sb = cell_sector[keyA[0]][keyA[1]], cell_sector[keyB[0]][keyB[1]]
                if kind == "open":
                    two_sided(keyA, keyB, p, q)
                elif kind == "wall":
                    one_sided(keyA, p, q, sa)
                    one_sided(keyB, q, p, sb)
                else:
                    dx, dy = (q[0] - p[0]) // 4, (q[1] - p[1]) // 4
                    m1 = (p[0] + dx, p[1] + dy)
                    m2 = (p[0] + 3 * dx, p[1] + 3 * dy)
                    one_sided(keyA, p, m1, sa)
                    two_sided(keyA, keyB, m1, m2)
                    one_sided(keyA, m2, q, sa)
                    one_sided(keyB, q, m2, sb)
                    one_sided(keyB, m1, p, sb)
            elif ra:
                one_sided(keyA, p, q, cell_sector[keyA[0]][keyA[1]])
            elif rb:
                one_sided(keyB, q, p, cell_sector[keyB[0]][keyB[1]])

        # vertical grid edges x = i*cs between cell (i-1,j) [west] and (i,j) [east];
        # going north (p->q) the east cell is on the right.
        for i in range(gx + 1):
            for j in range(gy):
                west = (i - 1, j) if i > 0 else None
                east = (i, j) if i < gx else None
                edge(east, west, corner(i, j), corner(i, j + 1))
        # horizontal grid edges y = j*cs between (i,j-1) [south] and (i,j) [north];
        # going east the south cell is on the right.
        for j in range(gy + 1):
            for i in range(gx):
                south = (i, j - 1) if j > 0 else None
                north = (i, j) if j < gy else None
                edge(south, north, corner(i, j), corner(i + 1, j))

        # 4. inner polygons
        inner: Dict[Tuple[int, int], dict] = {}
        start_cell = None
        room_cells = [k for k in cell_segs]
        for key in room_cells:
            if rng.below(100) < cfg.inner_pct:
                shape = rng.pick(_INNER_SHAPES)
                solid = rng.chance(1, 2)
                inner[key] = self._add_inner(key, shape, solid, cell_segs, corner)
        plain = [k for k in room_cells if k not in inner]
        start_cell = plain[rng.below(len(plain))] if plain else room_cells[0]
        # 5. BSP
        self.root_child = self._kd(0, gx, 0, gy, cell_segs, inner, corner)
        if not (self.root_child & 0x8000):
            assert self.root_child == len(self.nodes) - 1
        else:  # single subsector level: fabricate a root node is impossible; keep at least 2 rooms
            raise AssertionError("level needs at least two subsectors")
        # 6. things
        sx, sy = corner(*start_cell)
        self.things.append((sx + cs // 2, sy + cs // 2, rng.below(8) * 45, 1, 7))
        for key in plain[:6]:
            cx, cy = corner(*key)
            self.things.append((cx + cs // 3, cy + cs // 3, 0, 2035, 7))
        if cfg.thing_pct:
            # decorations: floor-standing, hanging, a type without metadata (9999) and one whose sprite lump is
            # missing (2014): the last two are skipped by the loaders (visitor.rs:1063-1094)
            kinds = [2035, 48, 34, 2028, 63, 46, 9999, 2014]
            for key in plain:
                if rng.below(100) >= cfg.thing_pct:
                    continue
                cx, cy = corner(*key)
                for _ in range(1 + rng.below(3)):
                    self.things.append((cx + cs // 2 + rng.below(129) - 64, cy + cs // 2 + rng.below(129) - 64,
                                        rng.below(8) * 45, rng.pick(kinds), 7))
        return self

    def _connected(self, room) -> bool:
        gx, gy = self.cfg.gx, self.cfg.gy
        cells = [(i, j) for i in range(gx) for j in range(gy) if room[i][j]]
        if len(cells) < 2:
            return False
        seen = {cells[0]}
        stack = [cells[0]]
        while stack:
            i, j = stack.pop()
            for a, b in ((i + 1, j), (i - 1, j), (i, j + 1), (i, j - 1)):
                if 0 <= a < gx and 0 <= b < gy and room[a][b] and (a, b) not in seen:
                    seen.add((a, b))
                    stack.append((a, b))
        return len(seen) == len(cells)

    def _spanning_tree(self, room):
        gx, gy = self.cfg.gx, self.cfg.gy
        cells = [(i, j) for i in range(gx) for j in range(gy) if room[i][j]]
        seen = {cells[0]}
        frontier = [cells[0]]
        tree = set()
        while frontier:
            k = self.rng.below(len(frontier))
            i, j = frontier[k]
            nbrs = [(a, b) for a, b in ((i + 1, j), (i - 1, j), (i, j + 1), (i, j - 1))
                    if 0 <= a < gx and 0 <= b < gy and room[a][b] and (a, b) not in seen]
            if not nbrs:
                frontier.pop(k)
                continue
            n = nbrs[self.rng.below(len(nbrs))]
            seen.add(n)
            tree.add(((i, j), n))
            frontier.append(n)
        return tree

    def _add_inner(self, key, shape, solid, cell_segs, corner):
        cfg, rng = self.cfg, self.rng
        cs = cfg.cell
        bx, by = corner(*key)
        pts = [(bx + x * cs // 256, by + y * cs // 256) for x, y in shape]
        outer_sector = self.cell_sector[key[0]][key[1]]
        so = self.sectors[outer_sector]
        inner_sector = -1
        if not solid:
            df = rng.pick([-24, -16, 8, 16, 24, 40, 72])
            dc = rng.pick([0, 0, -16, -32, 24])
            fl = so.floor + df
            ce = max(fl + 56, so.ceil + dc) if so.ceil_flat != "F_SKY1" else so.ceil
            self.sectors.append(Sector(fl, ce, self._nukage(rng.pick(FLOOR_FLATS)), so.ceil_flat if so.ceil_flat == "F_SKY1"
                                       else rng.pick(CEIL_FLATS), rng.pick([128, 160, 192, 255, 96])))
            inner_sector = len(self.sectors) - 1
        n = len(pts)
        edge_segs_outer = []
        inner_segs = []
        tex = rng.pick(WALL_TEX)
        for e in range(n):
            p, q = pts[e], pts[(e + 1) % n]      # CCW: cell (exterior) is on the right
            a, b = self.vert(*p), self.vert(*q)
            if solid:
                sd = self.side(outer_sector, middle=tex, xoff=e * 8)
                li = self.line(a, b, sd, -1, flags=0x0001 | (0x0010 if e % 3 == 0 else 0))
                edge_segs_outer.append(Seg(a, b, li, 0, 0))
            else:
                flags = 0x0004 | (0x0008 if e % 2 else 0) | (0x0010 if e % 3 == 0 else 0)
                s1 = self.side(outer_sector, upper=tex, lower=rng.pick(STEP_TEX), xoff=e * 4, yoff=0)
                s2 = self.side(inner_sector, upper=tex, lower=rng.pick(STEP_TEX))
                li = self.line(a, b, s1, s2, flags=flags)
                edge_segs_outer.append(Seg(a, b, li, 0, 0))
                inner_segs.append(Seg(b, a, li, 1, 0))
        return dict(pts=pts, solid=solid, outer=edge_segs_outer, inner=inner_segs)

    # -- BSP ------------------------------------------------------------------------
    def _emit_subsector(self, segs: List[Seg]) -> int:
        assert segs
        first = len(self.segs)
        self.segs.extend(segs)
        self.ssectors.append((len(segs), first))
        return 0x8000 | (len(self.ssectors) - 1)

    def _bbox_of_segs(self, segs: List[Seg]):
        xs = [self.vertices[s.v1][0] for s in segs] + [self.vertices[s.v2][0] for s in segs]
        ys = [self.vertices[s.v1][1] for s in segs] + [self.vertices[s.v2][1] for s in segs]
        return (max(ys), min(ys), min(xs), max(xs))   # top, bottom, left, right

    @staticmethod
    def _bbox_union(a, b):
        return (max(a[0], b[0]), min(a[1], b[1]), min(a[2], b[2]), max(a[3], b[3]))

    def _emit_node(self, x, y, dx, dy, rbox, lbox, rchild, lchild) -> int:
        self.nodes.append((x, y, dx, dy, rbox, lbox, rchild, lchild))
        return len(self.nodes) - 1

    def _cell_tree(self, key, cell_segs, inner):
        """Returns (child id, bbox) for one room cell."""
        segs = list(cell_segs[key])
        if key not in inner:
            return self._emit_subsector(segs), self._bbox_of_segs(segs)
        info = inner[key]
        pts = info["pts"]
        n = len(pts)
        remaining = segs + list(info["outer"])
        # cuts: for each polygon edge e (p->q, CCW) the *outer* side is the right side of p->q.
        pieces = []   # (line, outer seg list)
        ncuts = n - 1 if info["solid"] else n
        for e in range(n):
            p, q = pts[e], pts[(e + 1) % n]
            dx, dy = q[0] - p[0], q[1] - p[1]
            if e >= ncuts:
                break
            outer, keep = [], []
            for s in remaining:
                a, b = self.vertices[s.v1], self.vertices[s.v2]
                # side value > 0: left of p->q (inner side); < 0: right (outer side)
                sa = dx * (a[1] - p[1]) - dy * (a[0] - p[0])
                sb = dx * (b[1] - p[1]) - dy * (b[0] - p[0])
                if sa == 0 and sb == 0:
                    # collinear with the cut: it is polygon edge e itself (belongs to the outer piece)
                    outer.append(s)
                elif sa <= 0 and sb <= 0:
                    outer.append(s)
                elif sa >= 0 and sb >= 0:
                    keep.append(s)
                else:
                    # split at the intersection
                    t_num, t_den = sa, sa - sb
                    ix_num = a[0] * t_den + (b[0] - a[0]) * t_num
                    iy_num = a[1] * t_den + (b[1] - a[1]) * t_num
                    assert ix_num % t_den == 0 and iy_num % t_den == 0, "non-integer seg split"
                    m = self.vert(ix_num // t_den, iy_num // t_den)
                    ld = self.linedefs[s.linedef]
                    anchor = self.vertices[ld.v1] if s.direction == 0 else self.vertices[ld.v2]
                    pm = self.vertices[m]
                    off_m = _iround_hypot(pm[0] - anchor[0], pm[1] - anchor[1])
                    s1 = Seg(s.v1, m, s.linedef, s.direction, s.offset)
                    s2 = Seg(m, s.v2, s.linedef, s.direction, off_m)
                    (outer if sa < 0 else keep).append(s1)
                    (outer if sb < 0 else keep).append(s2)
            assert outer, "empty outer piece"
            pieces.append(((p[0], p[1], dx, dy), outer))
            remaining = keep
        if info["solid"]:
            last = remaining          # region beyond the final edge (convex)
        else:
            last = list(info["inner"])
            assert not remaining, "leftover segs after the final cut"
        assert last
        # build the chain bottom-up: node_k = (cut_k: right=outer piece k, left=rest)
        child = self._emit_subsector(last)
        box = self._bbox_of_segs(last)
        for (line, outer) in reversed(pieces):
            rchild = self._emit_subsector(outer)
            rbox = self._bbox_of_segs(outer)
            node = self._emit_node(line[0], line[1], line[2], line[3], rbox, box, rchild, child)
            child = node
            box = self._bbox_union(rbox, box)
        return child, box

    def _kd(self, i0, i1, j0, j1, cell_segs, inner, corner):
        cells = [(i, j) for i in range(i0, i1) for j in range(j0, j1) if self.room[i][j]]
        assert cells
        child, _ = self._kd_rec(i0, i1, j0, j1, cell_segs, inner, corner)
        return child

    def _kd_rec(self, i0, i1, j0, j1, cell_segs, inner, corner):
        cells = [(i, j) for i in range(i0, i1) for j in range(j0, j1) if self.room[i][j]]
        if len(cells) == 1:
            return self._cell_tree(cells[0], cell_segs, inner)
        # candidate splits: choose the one that balances room counts (both sides non-empty)
        best = None
        for axis in (0, 1):
            lo, hi = (i0, i1) if axis == 0 else (j0, j1)
            for k in range(lo + 1, hi):
                na = sum(1 for c in cells if c[axis] < k)
                nb = len(cells) - na
                if na == 0 or nb == 0:
                    continue
                score = (abs(na - nb), axis, k)
                if best is None or score < best[0]:
                    best = (score, axis, k)
        assert best is not None
        _, axis, k = best
        cs = self.cfg.cell
        if axis == 0:
            lo_child, lo_box = self._kd_rec(i0, k, j0, j1, cell_segs, inner, corner)
            hi_child, hi_box = self._kd_rec(k, i1, j0, j1, cell_segs, inner, corner)
            px, py = corner(k, j0)
            # partition pointing north: right side = east = high side
            node = self._emit_node(px, py, 0, (j1 - j0) * cs, hi_box, lo_box, hi_child, lo_child)
        else:
            lo_child, lo_box = self._kd_rec(i0, i1, j0, k, cell_segs, inner, corner)
            hi_child, hi_box = self._kd_rec(i0, i1, k, j1, cell_segs, inner, corner)
            px, py = corner(i0, k)
            # partition pointing east: right side = south = low side
            node = self._emit_node(px, py, (i1 - i0) * cs, 0, lo_box, hi_box, lo_child, hi_child)
        return node, self._bbox_union(lo_box, hi_box)

    # -- serialisation --------------------------------------------------------------
    def lumps(self) -> List[Tuple[str, bytes]]:
        things = b"".join(struct.pack("<hhhHH", *t) for t in self.things)
        lines = b"".join(struct.pack("<HHHHHhh", l.v1, l.v2, l.flags, l.special, l.tag, l.right, l.left)
                         for l in self.linedefs)
        sides = b"".join(struct.pack("<hh8s8s8sH", s.xoff, s.yoff, _name8(s.upper), _name8(s.lower),
                                     _name8(s.middle), s.sector) for s in self.sidedefs)
        verts = b"".join(struct.pack("<hh", *v) for v in self.vertices)
        segs = b""
        for s in self.segs:
            a, b = self.vertices[s.v1], self.vertices[s.v2]
            ang = int(round(math.atan2(b[1] - a[1], b[0] - a[0]) * 32768.0 / math.pi)) & 0xFFFF
            segs += struct.pack("<HHHHHH", s.v1, s.v2, ang, s.linedef, s.direction, s.offset & 0xFFFF)
        ssec = b"".join(struct.pack("<HH", n, f) for n, f in self.ssectors)
        nodes = b""
        for (x, y, dx, dy, rb, lb, rc, lc) in self.nodes:
            nodes += struct.pack("<hhhh4h4hHH", x, y, dx, dy, *rb, *lb, rc, lc)
        secs = b"".join(struct.pack("<hh8s8shHH", s.floor, s.ceil, _name8(s.floor_flat), _name8(s.ceil_flat),
                                    s.light, s.stype, s.tag) for s in self.sectors)
        nsec = len(self.sectors)
        reject = bytes((nsec * nsec + 7) // 8)
        blockmap = struct.pack("<hhHH", self.cfg.origin[0], self.cfg.origin[1], 0, 0)
        return [(self.name, b""), ("THINGS", things), ("LINEDEFS", lines), ("SIDEDEFS", sides),
                ("VERTEXES", verts), ("SEGS", segs), ("SSECTORS", ssec), ("NODES", nodes),
                ("SECTORS", secs), ("REJECT", reject), ("BLOCKMAP", blockmap)]


# --------------------------------------------------------------------------------------
# IWAD assembly
# --------------------------------------------------------------------------------------
def assemble_wad(lumps: Sequence[Tuple[str, bytes]], ident: bytes = b"IWAD") -> bytes:
    body = bytearray()
    directory = bytearray()
    pos = 12
    for name, data in lumps:
        directory += struct.pack("<ii8s", pos if data else 0, len(data), _name8(name))
        body += data
        pos += len(data)
    header = struct.pack("<4sii", ident, len(lumps), 12 + len(body))
    return bytes(header) + bytes(body) + bytes(directory)


def build_iwad(seed: int = 1, maps: Sequence[str] = ("E1M1",), cfg: Optional[SynthConfig] = None,
               map_seeds: Optional[Sequence[int]] = None) -> bytes:
    """Build a complete synthetic IWAD.  ``seed`` drives graphics; level ``k`` uses
    ``map_seeds[k]`` (default ``seed*100 + k``).  The default config is E1M1-scale
    (about 85-100 sectors, 450-500 linedefs, 700-800 segs, 230-260 subsectors)."""
    cfg = cfg or SynthConfig()
    rng = SplitMix64(seed)
    playpal = make_playpal()
    colormap = make_colormap(playpal)
    patches, tex, flats = make_graphics(rng, masked=cfg.mid_pct > 0, anim=cfg.anim, odd=cfg.odd_tex)
    pnames, texture1 = make_pnames_texture1(list(patches.keys()), tex)
    lumps: List[Tuple[str, bytes]] = [("PLAYPAL", playpal), ("COLORMAP", colormap)]
    for k, name in enumerate(maps):
        ms = map_seeds[k] if map_seeds is not None else seed * 100 + k
        lumps += LevelBuilder(name, ms, cfg).generate().lumps()
    lumps += [("TEXTURE1", texture1), ("PNAMES", pnames)]
    lumps += [("P_START", b"")] + [(n, d) for n, d in patches.items()] + [("P_END", b"")]
    sprites = [("PLAYA1", encode_picture(_img_gradient(rng, 16, 32, 2), 8, 30))]
    if cfg.thing_pct:
        # <prefix><frame>0 = rotation-less, <prefix><frame>1 = first rotation (visitor.rs:1070-1088 tries 0 then 1)
        for name, w, h, hue in (("BAR1A0", 24, 32, 6), ("ELECA0", 32, 120, 8), ("CANDA0", 10, 16, 7),
                                ("COLUA0", 24, 48, 5), ("GOR1A0", 24, 64, 2), ("TREDA1", 26, 80, 12)):
            sprites.append((name, encode_picture(_img_sprite(rng, w, h, hue), w // 2, h - 4)))
    lumps += [("S_START", b"")] + sprites + [("S_END", b"")]
    lumps += [("F_START", b"")] + [(n, d) for n, d in flats.items()] + [("F_END", b"")]
    return assemble_wad(lumps)


if __name__ == "__main__":
    import sys
    out = sys.argv[1] if len(sys.argv) > 1 else "synth.wad"
    data = build_iwad(1, E1_MAPS[:1])
    open(out, "wb").write(data)
    print(out, len(data), "bytes")
