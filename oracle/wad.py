"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's `wad/` crate.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product path (rust-doom_b200/, libb2d.so) never does.

Parity status: **unpinned** for everything except WadName -- the reference holds no golden
vectors, fixtures or WAD for this path; the 15 `test_wad_name`
assertions (wad/src/name.rs:168-189) are ported in tests/test_oracle_wad.py.

Each function cites the reference file:line it follows.  numpy is used for record decoding;
pure-Python loops only for picture posts (small).
"""
from __future__ import annotations

import struct
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import numpy as np


class WadError(Exception):
    """Mirrors wad::ErrorKind::{CorruptWad, Io} (wad/src/errors.rs:9-19)."""


# ---------------------------------------------------------------------------------------------
# names  (wad/src/name.rs:41-75)
# ---------------------------------------------------------------------------------------------
_VALID = set(b"ABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789_-[]%\\")


def wad_name(value: bytes) -> bytes:
    """WadName::from_bytes: upper-case, stop at NUL, reject other bytes, max 8 unless NUL-terminated
    earlier.  Returns the 8-byte zero-padded name."""
    name = bytearray(8)
    nulled = False
    for i, src in enumerate(value[:8]):
        if src >= 0x80:
            raise WadError("invalid byte %#x in wad name %r" % (src, value))
        up = src - 32 if 97 <= src <= 122 else src
        if up == 0:
            nulled = True
            break
        if up not in _VALID:
            raise WadError("invalid byte %#x in wad name %r" % (src, value))
        name[i] = up
    if not (nulled or len(value) <= 8):
        raise WadError("wad name too long %r" % (value,))
    return bytes(name)


def is_untextured(name: bytes) -> bool:      # wad/src/util.rs:4-6
    return name[0:1] == b"-" and name[1:2] == b"\0"


def is_sky_flat(name: bytes) -> bool:        # wad/src/util.rs:8-10
    return name == b"F_SKY1\0\0"


# ---------------------------------------------------------------------------------------------
# record layouts  (wad/src/types.rs:20-169)
# ---------------------------------------------------------------------------------------------
THING = np.dtype([("x", "<i2"), ("y", "<i2"), ("angle", "<i2"), ("type", "<u2"), ("flags", "<u2")])
VERTEX = np.dtype([("x", "<i2"), ("y", "<i2")])
LINEDEF = np.dtype([("v1", "<u2"), ("v2", "<u2"), ("flags", "<u2"), ("special", "<u2"),
                    ("tag", "<u2"), ("right", "<i2"), ("left", "<i2")])
SIDEDEF = np.dtype([("xoff", "<i2"), ("yoff", "<i2"), ("upper", "S8"), ("lower", "S8"),
                    ("middle", "S8"), ("sector", "<u2")])
SECTOR = np.dtype([("floor", "<i2"), ("ceil", "<i2"), ("floor_tex", "S8"), ("ceil_tex", "S8"),
                   ("light", "<i2"), ("type", "<u2"), ("tag", "<u2")])
SUBSECTOR = np.dtype([("num_segs", "<u2"), ("first_seg", "<u2")])
SEG = np.dtype([("v1", "<u2"), ("v2", "<u2"), ("angle", "<u2"), ("linedef", "<u2"),
                ("direction", "<u2"), ("offset", "<u2")])
NODE = np.dtype([("x", "<i2"), ("y", "<i2"), ("dx", "<i2"), ("dy", "<i2"),
                 ("rbox", "<i2", (4,)), ("lbox", "<i2", (4,)), ("right", "<u2"), ("left", "<u2")])
assert (THING.itemsize, VERTEX.itemsize, LINEDEF.itemsize, SIDEDEF.itemsize, SECTOR.itemsize,
        SUBSECTOR.itemsize, SEG.itemsize, NODE.itemsize) == (10, 4, 14, 30, 26, 4, 12, 28)


def _raw_name(b: bytes) -> bytes:
    """numpy 'S8' strips trailing NULs; restore the padded, validated form."""
    return wad_name(b)


# ---------------------------------------------------------------------------------------------
# archive  (wad/src/archive.rs:36-106, 172-242)
# ---------------------------------------------------------------------------------------------
class Archive:
    def __init__(self, data: bytes, overlays=()):
        """`data` is the IWAD (all the reference opens, archive.rs:69-72).  `overlays`: PWAD files applied in order, Doom-engine
        style -- their lumps are appended to the directory (a later lump of a name wins every by-name lookup), a level of
        an existing name replaces that level in place, new level names are appended, and flats / sprites between a PWAD's
        FF_START..FF_END / SS_START..SS_END (or F_START / S_START) markers are added to the IWAD's."""
        self.data = data
        self.files: List[bytes] = []
        self.file_ranges: List[Tuple[int, int]] = []
        self.lumps: List[Tuple[bytes, int, int]] = []          # (name, offset, size)
        self.lump_file: List[int] = []
        self.index_map: Dict[bytes, int] = {}                  # later duplicate wins (archive.rs:85)
        self.levels: List[int] = []
        for k, blob in enumerate((data,) + tuple(overlays)):
            self._add_file(blob, k == 0)

    def _add_file(self, data: bytes, iwad: bool) -> None:
        if len(data) < 12:
            raise WadError("bad wad header")
        ident, num_lumps, info_off = struct.unpack_from("<4sii", data, 0)
        if ident != (b"IWAD" if iwad else b"PWAD"):            # archive.rs:69-72
            raise WadError("bad wad header identifier %r" % ident)
        if info_off < 0 or info_off + 16 * num_lumps > len(data) or num_lumps < 0:
            raise WadError("bad lump info table")
        fi = len(self.files)
        self.files.append(data)
        first = len(self.lumps)
        for i in range(num_lumps):
            pos, size, raw = struct.unpack_from("<ii8s", data, info_off + 16 * i)
            name = wad_name(raw)                               # invalid byte fails the open
            self.index_map[name] = len(self.lumps)
            self.lumps.append((name, pos, size))
            self.lump_file.append(fi)
            if name == b"THINGS\0\0":                          # archive.rs:92-97
                if i == 0:
                    raise WadError("THINGS lump without a level marker")
                marker = first + i - 1
                same = [k for k, lv in enumerate(self.levels) if self.lumps[lv][0] == self.lumps[marker][0]]
                if same:
                    for k in same:
                        self.levels[k] = marker
                else:
                    self.levels.append(marker)
        self.file_ranges.append((first, len(self.lumps)))

    def marker_ranges(self, starts, ends) -> List[Tuple[int, int]]:
        """[first, last) lump ranges between the start / end markers of every file that has them (the IWAD must)."""
        starts = [wad_name(x) for x in starts]
        ends = [wad_name(x) for x in ends]
        out = []
        for f, (lo, hi) in enumerate(self.file_ranges):
            a = b = -1
            for i in range(lo, hi):
                if self.lumps[i][0] in starts:
                    a = i
                if self.lumps[i][0] in ends:
                    b = i
            if f == 0 and (a < 0 or b < 0):
                raise WadError("missing required lump %r" % (starts[0] if a < 0 else ends[0]))
            if a >= 0 and b >= 0:
                out.append((a, b))
        return out

    @classmethod
    def open(cls, path: str) -> "Archive":
        try:
            with open(path, "rb") as f:
                return cls(f.read())
        except OSError as e:
            raise WadError("io: %s" % e)

    def num_levels(self) -> int:
        return len(self.levels)

    def level_name(self, level_index: int) -> bytes:
        return self.lumps[self.levels[level_index]][0]

    def named(self, name: bytes) -> Optional[int]:
        return self.index_map.get(wad_name(name))

    def required(self, name: bytes) -> int:
        idx = self.named(name)
        if idx is None:
            raise WadError("missing required lump %r" % name)
        return idx

    def read(self, index: int) -> bytes:
        if not (0 <= index < len(self.lumps)):
            raise WadError("missing required lump index %d" % index)
        _, pos, size = self.lumps[index]
        if size == 0:
            return b""
        data = self.files[self.lump_file[index]]
        if pos < 0 or size < 0 or pos + size > len(data):        # i32 -> usize wrap makes the read fail
            raise WadError("lump %d out of file bounds" % index)
        return data[pos:pos + size]

    def decode_vec(self, index: int, dtype: np.dtype) -> np.ndarray:
        """LumpReader::decode_vec: size > 0 and a multiple of the element size (archive.rs:172-190)."""
        buf = self.read(index)
        if len(buf) == 0 or len(buf) % dtype.itemsize != 0:
            raise WadError("bad lump size %d for element %d (lump %d %r)" %
                           (len(buf), dtype.itemsize, index, self.lumps[index][0]))
        return np.frombuffer(buf, dtype=dtype)

    def read_blobs(self, index: int, blob: int) -> List[bytes]:
        buf = self.read(index)
        if len(buf) == 0 or len(buf) % blob != 0:
            raise WadError("bad blob lump size %d" % len(buf))
        return [buf[i:i + blob] for i in range(0, len(buf), blob)]


# ---------------------------------------------------------------------------------------------
# level  (wad/src/level.rs:13-81)
# ---------------------------------------------------------------------------------------------
class Level:
    def __init__(self, wad: Archive, level_index: int):
        if not (0 <= level_index < wad.num_levels()):
            raise WadError("no such level %d" % level_index)
        start = wad.levels[level_index]
        self.name = wad.lumps[start][0]
        self.things = wad.decode_vec(start + 1, THING)
        self.linedefs = wad.decode_vec(start + 2, LINEDEF)
        self.sidedefs = wad.decode_vec(start + 3, SIDEDEF)
        self.vertices = wad.decode_vec(start + 4, VERTEX)
        self.segs = wad.decode_vec(start + 5, SEG)
        self.subsectors = wad.decode_vec(start + 6, SUBSECTOR)
        self.nodes = wad.decode_vec(start + 7, NODE)
        self.sectors = wad.decode_vec(start + 8, SECTOR)
        # name validation happens at deserialisation time in the reference (name.rs:132-139)
        for arr, fields in ((self.sidedefs, ("upper", "lower", "middle")),
                            (self.sectors, ("floor_tex", "ceil_tex"))):
            raw = arr.tobytes()
            for f in fields:
                off = arr.dtype.fields[f][1]
                for i in range(len(arr)):
                    wad_name(raw[i * arr.dtype.itemsize + off:i * arr.dtype.itemsize + off + 8])

    # navigation: wad/src/level.rs:83-161
    def seg_sidedef_index(self, seg) -> int:
        line = self.linedefs[seg["linedef"]] if seg["linedef"] < len(self.linedefs) else None
        if line is None:
            return -1
        idx = int(line["right"] if seg["direction"] == 0 else line["left"])
        return idx if 0 <= idx < len(self.sidedefs) else -1

    def seg_back_sidedef_index(self, seg) -> int:
        line = self.linedefs[seg["linedef"]] if seg["linedef"] < len(self.linedefs) else None
        if line is None:
            return -1
        idx = int(line["right"] if seg["direction"] == 1 else line["left"])
        return idx if 0 <= idx < len(self.sidedefs) else -1

    def sector_min_light(self, sector_id: int) -> int:
        """level.rs:163-182: min light over sectors adjacent through two-sided lines."""
        m = int(self.sectors[sector_id]["light"])
        for l in self.linedefs:
            r, le = int(l["right"]), int(l["left"])
            if r < 0 or le < 0 or r >= len(self.sidedefs) or le >= len(self.sidedefs):
                continue
            a, b = int(self.sidedefs[r]["sector"]), int(self.sidedefs[le]["sector"])
            if a == sector_id and b != sector_id and b < len(self.sectors):
                m = min(m, int(self.sectors[b]["light"]))
            elif b == sector_id and a != sector_id and a < len(self.sectors):
                m = min(m, int(self.sectors[a]["light"]))
        return m


# ---------------------------------------------------------------------------------------------
# pictures  (wad/src/image.rs:39-252)
# ---------------------------------------------------------------------------------------------
MAX_IMAGE_SIZE = 4096
TRANSPARENT_NEW = 0xFF00      # Image::new fill (image.rs:30)
TRANSPARENT_DECODED = 0xFFFF  # from_buffer fill (image.rs:63)


def decode_picture(buf: bytes) -> Tuple[np.ndarray, int, int]:
    """Image::from_buffer -> (row-major u16 pixels [h,w], x_offset, y_offset)."""
    if len(buf) < 8:
        raise WadError("image missing header")
    w, h, xo, yo = struct.unpack_from("<HHhh", buf, 0)
    if w > MAX_IMAGE_SIZE or h > MAX_IMAGE_SIZE:
        raise WadError("image too large %dx%d" % (w, h))
    if len(buf) < 8 + 4 * w:
        raise WadError("unfinished image column directory")
    px = np.full((h, w), TRANSPARENT_DECODED, dtype=np.uint16)
    offs = struct.unpack_from("<%dI" % w, buf, 8)
    n = len(buf)
    for x in range(w):
        p = offs[x]
        if p >= n:
            raise WadError("invalid image column offset in %d" % x)
        last_row = -1
        while True:
            if p >= n:
                raise WadError("unfinished image column %d" % x)
            row = buf[p]
            p += 1
            if row == 255:
                break
            # tall patches (DeePsea convention): a post whose topdelta does not exceed the previous post's is relative to
            # it; stock patches have strictly increasing posts, so this changes nothing the reference can represent
            if last_row >= 0 and row <= last_row:
                row += last_row
            last_row = row
            if p >= n:
                raise WadError("missing image run length")
            ln = buf[p]
            p += 1
            if row + ln > h:
                raise WadError("image run too big: column %d" % x)
            if p >= n:
                raise WadError("image missing padding byte 1")
            p += 1
            if n - p < ln:
                raise WadError("image source underrun")
            if ln:
                px[row:row + ln, x] = np.frombuffer(buf, dtype=np.uint8, count=ln, offset=p)
            p += ln
            if p >= n:
                raise WadError("image missing padding byte 2")
            p += 1
    return px, xo, yo


def blit(dest: np.ndarray, src: np.ndarray, ox: int, oy: int, ignore_transparency: bool) -> None:
    """Image::blit (image.rs:171-252): clipped copy; masked copy keeps dest where src bit15 set."""
    dh, dw = dest.shape
    sh, sw = src.shape
    if ox >= dw or oy >= dh:
        return
    y0 = -oy if oy < 0 else 0
    x0 = -ox if ox < 0 else 0
    y1 = sh if dh > sh + oy else dh - oy
    x1 = sw if dw > sw + ox else dw - ox
    if x1 <= x0 or y1 <= y0:
        return          # fully off the top/left: the reference would panic on the slice; we skip
    s = src[y0:y1, x0:x1]
    d = dest[y0 + oy:y1 + oy, x0 + ox:x1 + ox]
    if ignore_transparency:
        d[...] = s
    else:
        mask = (s >> 15) != 0
        d[...] = np.where(mask, d, s)


# ---------------------------------------------------------------------------------------------
# texture directory  (wad/src/tex.rs:53-107, 358-410, 499-606)
# ---------------------------------------------------------------------------------------------
class TextureDirectory:
    def __init__(self, wad: Archive):
        self.palettes = wad.read_blobs(wad.required(b"PLAYPAL"), 768)
        self.colormaps = wad.read_blobs(wad.required(b"COLORMAP"), 256)
        self.patches: List[Tuple[bytes, Optional[np.ndarray]]] = self._read_patches(wad)
        self.textures: "OrderedDict[bytes, np.ndarray]" = OrderedDict()
        for lump_name in (b"TEXTURE1", b"TEXTURE2"):
            idx = wad.named(lump_name)
            if idx is None:
                continue
            self._read_textures(wad.read(idx))
        self.flats: "OrderedDict[bytes, bytes]" = OrderedDict()
        for start, end in wad.marker_ranges((b"F_START", b"FF_START"), (b"F_END", b"FF_END")):
            for i in range(start, end):
                name, _, size = wad.lumps[i]
                if size == 0:
                    continue
                self.flats[name] = wad.read(i)          # a PWAD's flat of the same name wins
        # sprites share the texture map and may shadow a texture name (tex.rs:475-497)
        for s0, s1 in wad.marker_ranges((b"S_START", b"SS_START"), (b"S_END", b"SS_END")):
            for i in range(s0 + 1, s1):
                buf = wad.read(i)                       # a read failure fails the load (tex.rs:484-485)
                try:
                    px, _, _ = decode_picture(buf)
                except WadError:
                    continue                            # a decode failure skips the sprite (tex.rs:486-493)
                self.textures[wad.lumps[i][0]] = px

    @staticmethod
    def _read_patches(wad: Archive):
        buf = wad.read(wad.required(b"PNAMES"))
        if len(buf) < 4:
            raise WadError("missing number of patches in PNAMES")
        (n,) = struct.unpack_from("<I", buf, 0)
        out = []
        for i in range(n):
            raw = buf[4 + 8 * i:12 + 8 * i]
            if len(raw) < 8:
                continue
            try:
                name = wad_name(raw)
            except WadError:
                continue
            idx = wad.named(name)
            if idx is None:
                out.append((name, None))
                continue
            pbuf = wad.read(idx)                    # read errors propagate (tex.rs:384-385)
            try:
                px, _, _ = decode_picture(pbuf)
                out.append((name, px))
            except WadError:
                out.append((name, None))
        return out

    def _read_textures(self, lump: bytes) -> None:
        if len(lump) < 4:
            raise WadError("missing number of textures")
        (n,) = struct.unpack_from("<I", lump, 0)
        if 4 * n >= len(lump) - 4:                              # tex.rs:510-517
            raise WadError("textures lump too small for offsets")
        for i in range(n):
            (off,) = struct.unpack_from("<I", lump, 4 + 4 * i)
            if off >= len(lump):
                raise WadError("textures lump too small for offsets")
            if off + 22 > len(lump):
                continue
            raw, _masked, w, h, _cd, npatches = struct.unpack_from("<8sIHHIH", lump, off)
            try:
                name = wad_name(raw)
            except WadError:
                continue
            if w > MAX_IMAGE_SIZE or h > MAX_IMAGE_SIZE:
                continue
            img = np.full((h, w), TRANSPARENT_NEW, dtype=np.uint16)
            p = off + 22
            for k in range(npatches):
                if p + 10 > len(lump):
                    break
                ox, oy, pidx, _sd, _cm = struct.unpack_from("<hhHHH", lump, p)
                p += 10
                if oy <= 0:                                   # tex.rs:560-567
                    oy = 0
                if pidx < len(self.patches) and self.patches[pidx][1] is not None:
                    blit(img, self.patches[pidx][1], ox, oy, k == 0)
            self.textures[name] = img


# ---------------------------------------------------------------------------------------------
# light  (wad/src/light.rs:27-115, game/src/lights.rs:14-30)
# ---------------------------------------------------------------------------------------------
EFFECT_TYPES = (1, 2, 4, 13, 3, 12, 8, 17)   # light.rs:127-134


def light_byte(light: int, contrast: int) -> int:
    """The u8 the reference uploads for a static light: float32 arithmetic restated with numpy
    float32 (light.rs:113-115 `(level >> 3) / 31`, :82-91 contrast +-2/31 and clamp,
    lights.rs:26-29 `(clamp(level) * 255.0) as u8`)."""
    f = np.float32
    level = f(np.int16(light) >> 3) / f(31.0)
    if contrast:
        level = level + (f(2.0) / f(31.0) if contrast > 0 else f(-2.0) / f(31.0))
        level = f(1.0) if level > f(1.0) else (f(0.0) if level < f(0.0) else level)
    level = f(1.0) if level > f(1.0) else (f(0.0) if level < f(0.0) else level)
    v = float(level * f(255.0))
    return int(v) & 0xFF if v >= 0 else 0
